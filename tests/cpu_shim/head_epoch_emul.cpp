// head_epoch_emul.cpp -- runs one training epoch of the adaptive head twice ON THE CPU (tests/cpu_shim/cuda_shim.h):
//   (a) the launch-per-kernel sequence of csrc/head.cu (ac_head_train_epoch, option "head_fused" = 0), kernel by kernel,
//   (b) fused::head_epoch_kernel as a cooperative launch of G blocks,
// from identical states, and demands bit-identical parameters, AdamW moments and accumulated loss.  TEST INFRASTRUCTURE:
// the fused kernel was written without GPU access; this checks its phase structure, virtual-block arithmetic, barriers
// (the fiber scheduler shuffles the thread order, so a missing barrier changes the result) and operation order.
//   usage: head_epoch_emul D H0 H1 C n batch G loss(0 ce|1 bce) dropout_p ewc(0|1) seed
#include "cuda_shim.h"
#include "../../include/adaptive_b200.h"

#define SHIM_MAX_BLOCKS 16
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
namespace ac {
static inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
static inline float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
}  // namespace ac

#include "_gen_head_device.inc"

using namespace ac;

struct Head {
    ac_head_params p;
    std::vector<std::vector<float>> store;
};
static Head make_head(int D, int H0, int H1, int C, uint32_t seed, float scale) {
    Head h;
    h.p.D = D; h.p.H0 = H0; h.p.H1 = H1; h.p.C = C;
    const size_t sz[6] = {size_t(H0) * D, size_t(H0), size_t(H1) * H0, size_t(H1), size_t(C) * H1, size_t(C)};
    h.store.resize(6);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (int t = 0; t < 6; ++t) {
        h.store[t].resize(sz[t]);
        for (auto &x : h.store[t]) x = scale == 0.f ? 0.f : u(rng) * scale;
    }
    h.p.W0 = h.store[0].data(); h.p.b0 = h.store[1].data(); h.p.W1 = h.store[2].data();
    h.p.b1 = h.store[3].data(); h.p.W2 = h.store[4].data(); h.p.b2 = h.store[5].data();
    return h;
}
static Head clone(const Head &a) {
    Head h;
    h.p = a.p;
    h.store = a.store;
    h.p.W0 = h.store[0].data(); h.p.b0 = h.store[1].data(); h.p.W1 = h.store[2].data();
    h.p.b1 = h.store[3].data(); h.p.W2 = h.store[4].data(); h.p.b2 = h.store[5].data();
    return h;
}

struct Cfg {
    float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, wd = 0.01f, max_norm = 1.f, dropout_p = 0.1f, ewc_lambda = 100.f;
    int loss_kind = AC_LOSS_CE, first_step = 3, use_ewc = 0, ewc_C_old = 0;
    uint64_t seed = 11;
};

// ---- (a) csrc/head.cu: rowdot(), colacc(), sgemm(), fwd_bwd(), train_step_impl(), ac_head_train_epoch() restated as shim launches
static void k_rowdot(const float *X, const float *W, float *Y, int M, int N, int K, SgemmEpi epi) {
    shim::launch(dim3((N + RD_WARPS - 1) / RD_WARPS, (M + 31) / 32), dim3(RD_WARPS * 32), [&] { rowdot_kernel<1>(X, W, Y, M, N, K, epi); });
}
static void k_colacc(const float *G, const float *W, float *Z, int M, int R, int J, SgemmEpi epi) {
    shim::launch(dim3((J + 31) / 32, (M + 31) / 32), dim3(CA_GROUPS * 32), [&] { colacc_kernel(G, W, Z, M, R, J, epi); });
}
static void k_sgemm(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, float *C, int64_t ldc, int M,
                    int N, int K) {
    SgemmEpi none{EPI_NONE, nullptr, nullptr, nullptr};
    shim::launch(dim3((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM), dim3(256), [&] { sgemm_kernel(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, none); });
}

static void ref_step(const float *X, const void *targets, int B, Head &P, Head &Mo, Head &V, const Cfg &cfg, int step, float *stats,
                     TrainWs &w, const Head *fisher, const Head *star) {
    const ac_head_params *p = &P.p;
    const int D = p->D, H0 = p->H0, H1 = p->H1, C = p->C;
    const float *mask0 = nullptr, *mask1 = nullptr;
    if (cfg.dropout_p > 0.f) {
        const int64_t n0 = int64_t(B) * H0, n1 = int64_t(B) * H1;
        shim::launch(dim3(unsigned((n0 + 255) / 256)), dim3(256), [&] { dropout_mask_kernel(w.mask0, n0, cfg.dropout_p, cfg.seed, 2ull * step); });
        shim::launch(dim3(unsigned((n1 + 255) / 256)), dim3(256), [&] { dropout_mask_kernel(w.mask1, n1, cfg.dropout_p, cfg.seed, 2ull * step + 1); });
        mask0 = w.mask0; mask1 = w.mask1;
    }
    // fwd_bwd
    k_rowdot(X, p->W0, w.h0, B, H0, D, SgemmEpi{EPI_BIAS_RELU_MASK, p->b0, mask0, nullptr});
    k_rowdot(w.h0, p->W1, w.h1, B, H1, H0, SgemmEpi{EPI_BIAS_RELU_MASK, p->b1, mask1, nullptr});
    k_rowdot(w.h1, p->W2, w.z, B, C, H1, SgemmEpi{EPI_BIAS, p->b2, nullptr, nullptr});
    shim::launch(dim3((B + 3) / 4), dim3(128), [&] { loss_grad_kernel(w.z, targets, B, C, cfg.loss_kind, w.dz, w.row_loss); });
    shim::launch(dim3(1), dim3(32), [&] { reduce_loss_kernel(w.row_loss, B, stats + 0); });
    k_sgemm(w.dz, 1, C, w.h1, H1, 1, w.g.W2, H1, C, H1, B);
    shim::launch(dim3((C + 127) / 128), dim3(128), [&] { colsum_kernel(w.dz, B, C, w.g.b2); });
    k_colacc(w.dz, p->W2, w.dh1, B, C, H1, SgemmEpi{EPI_RELUGRAD_MASK, nullptr, mask1, w.h1});
    k_sgemm(w.dh1, 1, H1, w.h0, H0, 1, w.g.W1, H0, H1, H0, B);
    shim::launch(dim3((H1 + 127) / 128), dim3(128), [&] { colsum_kernel(w.dh1, B, H1, w.g.b1); });
    k_colacc(w.dh1, p->W1, w.dh0, B, H1, H0, SgemmEpi{EPI_RELUGRAD_MASK, nullptr, mask0, w.h0});
    k_sgemm(w.dh0, 1, H0, X, D, 1, w.g.W0, D, H0, D, B);
    shim::launch(dim3((H0 + 127) / 128), dim3(128), [&] { colsum_kernel(w.dh0, B, H0, w.g.b0); });
    // train_step_impl tail
    Flat6 g = flat_of(&w.g);
    if (cfg.use_ewc) {
        Flat6 lim = ewc_limits(p, cfg.ewc_C_old);
        const float scale = cfg.ewc_lambda / float(B);
        shim::launch(dim3(RED_BLOCKS), dim3(256), [&] { ewc_grad_penalty_kernel(flat_of(p), flat_of(&fisher->p), flat_of(&star->p), g, lim, 2.f * scale, w.partial, 1); });
        shim::launch(dim3(1), dim3(32), [&] { finalize_kernel(w.partial, RED_BLOCKS, scale, 0, stats + 1); });
    } else {
        stats[1] = 0.f;
    }
    shim::launch(dim3(RED_BLOCKS), dim3(256), [&] { sumsq_kernel(g, w.partial); });
    shim::launch(dim3(1), dim3(32), [&] { finalize_kernel(w.partial, RED_BLOCKS, 1.f, 1, stats + 2); });
    const float bc1 = 1.f - powf(cfg.beta1, float(step));
    const float bc2 = 1.f - powf(cfg.beta2, float(step));
    shim::launch(dim3(RED_BLOCKS * 2), dim3(256), [&] {
        adamw_kernel(flat_of(p), g, flat_of(&Mo.p), flat_of(&V.p), stats + 2, cfg.lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.wd, cfg.max_norm, bc1, sqrtf(bc2));
    });
}

int main(int argc, char **argv) {
    if (argc < 12) { printf("usage: %s D H0 H1 C n batch G loss dropout ewc seed\n", argv[0]); return 64; }
    const int D = atoi(argv[1]), H0 = atoi(argv[2]), H1 = atoi(argv[3]), C = atoi(argv[4]), n = atoi(argv[5]), batch = atoi(argv[6]);
    const int G = atoi(argv[7]);
    Cfg cfg;
    cfg.loss_kind = atoi(argv[8]);
    cfg.dropout_p = float(atof(argv[9]));
    cfg.use_ewc = atoi(argv[10]);
    cfg.ewc_C_old = cfg.use_ewc ? C - 1 : 0;             // the head "grew" by one class
    const uint32_t seed = uint32_t(atoi(argv[11]));
    if (G > SHIM_MAX_BLOCKS) return 64;

    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<float> X(size_t(n) * D);
    for (auto &x : X) x = u(rng);
    std::vector<int64_t> y(n), perm(n);
    std::vector<float> yf(size_t(n) * C);
    for (int i = 0; i < n; ++i) { y[i] = rng() % C; perm[i] = i; }
    for (auto &t : yf) t = (rng() % 10) < 3 ? 1.f : 0.f;
    std::shuffle(perm.begin(), perm.end(), rng);
    const void *targets = cfg.loss_kind == AC_LOSS_CE ? static_cast<const void *>(y.data()) : static_cast<const void *>(yf.data());

    Head Pa = make_head(D, H0, H1, C, seed + 1, 0.2f), Pb = clone(Pa);
    Head Ma = make_head(D, H0, H1, C, 0, 0.f), Mb = clone(Ma), Va = clone(Ma), Vb = clone(Ma);
    Head fisher = make_head(D, H0, H1, C, seed + 2, 1.f), star = make_head(D, H0, H1, C, seed + 3, 0.2f);
    for (auto &t : fisher.store) for (auto &x : t) x = fabsf(x);

    // workspaces (carve() of head.cu) for both paths
    TrainWs wa, wb;
    const size_t wsb = carve(wa, nullptr, batch, &Pa.p);
    std::vector<uint8_t> wsa(wsb + 256), wsbuf(wsb + 256);
    carve(wa, wsa.data(), batch, &Pa.p);
    carve(wb, wsbuf.data(), batch, &Pb.p);
    std::vector<float> xb_a(size_t(batch) * D), xb_b(size_t(batch) * D);
    std::vector<uint8_t> yb_a(size_t(batch) * std::max(C, 2) * 4 + 64), yb_b(size_t(batch) * std::max(C, 2) * 4 + 64);
    float stats_a[4] = {0, 0, 0, 0}, stats_b[4] = {0, 0, 0, 0}, acc_a = 0.f, acc_b = 0.f;

    // ---- (a) launch-per-kernel epoch
    int step = cfg.first_step;
    for (int off = 0; off < n; off += batch, ++step) {
        const int nb = std::min(batch, n - off);
        shim::launch(dim3(nb), dim3(128), [&] { gather_batch_kernel(X.data(), targets, perm.data() + off, nb, D, C, cfg.loss_kind, xb_a.data(), yb_a.data()); });
        ref_step(xb_a.data(), yb_a.data(), nb, Pa, Ma, Va, cfg, step, stats_a, wa, &fisher, &star);
        shim::launch(dim3(1), dim3(32), [&] { accum_loss_kernel(stats_a, &acc_a); });
    }

    // ---- (b) fused cooperative epoch
    const int steps = (n + batch - 1) / batch;
    std::vector<float2> bc(steps);
    for (int i = 0; i < steps; ++i)
        bc[i] = make_float2(1.f - powf(cfg.beta1, float(cfg.first_step + i)), sqrtf(1.f - powf(cfg.beta2, float(cfg.first_step + i))));
    std::vector<float> partial_ewc(RED_BLOCKS);
    fused::EpochArgs a{};
    a.X = X.data(); a.targets = targets; a.perm = perm.data(); a.n = n; a.batch = batch; a.first_step = cfg.first_step;
    a.p = Pb.p; a.m = Mb.p; a.v = Vb.p; a.w = wb; a.xb = xb_b.data(); a.yb = yb_b.data(); a.stats = stats_b; a.loss_accum = &acc_b;
    a.partial_ewc = partial_ewc.data(); a.bias_corr = bc.data();
    a.lr = cfg.lr; a.beta1 = cfg.beta1; a.beta2 = cfg.beta2; a.eps = cfg.eps; a.weight_decay = cfg.wd; a.max_norm = cfg.max_norm;
    a.dropout_p = cfg.dropout_p; a.loss_kind = cfg.loss_kind; a.seed = cfg.seed; a.use_ewc = cfg.use_ewc; a.ewc_C_old = cfg.ewc_C_old;
    a.ewc_lambda = cfg.ewc_lambda;
    if (cfg.use_ewc) { a.fisher = fisher.p; a.star = star.p; }
    shim::launch_cooperative(dim3(G), dim3(fused::FT), [&] { fused::head_epoch_kernel(a); }, seed + 99);

    // ---- compare
    long long bad = 0;
    const char *names[6] = {"W0", "b0", "W1", "b1", "W2", "b2"};
    for (int t = 0; t < 6; ++t) {
        long long d = 0;
        for (size_t i = 0; i < Pa.store[t].size(); ++i) {
            d += memcmp(&Pa.store[t][i], &Pb.store[t][i], 4) != 0;
            d += memcmp(&Ma.store[t][i], &Mb.store[t][i], 4) != 0;
            d += memcmp(&Va.store[t][i], &Vb.store[t][i], 4) != 0;
        }
        if (d) printf("  %s: %lld differing words (params + moments)\n", names[t], d);
        bad += d;
    }
    if (memcmp(&acc_a, &acc_b, 4) != 0) { printf("  loss accumulators differ: %.9g vs %.9g\n", acc_a, acc_b); ++bad; }
    double chk = 0;
    for (auto &t : Pa.store) for (float x : t) chk += x;
    printf("head_epoch_emul D=%d H0=%d H1=%d C=%d n=%d batch=%d G=%d loss=%d dropout=%.2f ewc=%d: %d steps, loss sum %.6f, param checksum %.6f -> %s\n",
           D, H0, H1, C, n, batch, G, cfg.loss_kind, cfg.dropout_p, cfg.use_ewc, steps, acc_a, chk, bad == 0 ? "FUSED == LAUNCH-PER-KERNEL" : "MISMATCH");
    return bad == 0 ? 0 : 1;
}
