// CPU emulation of pbs_kernel_mb2_wide (fhe_icp_b200/csrc/pbs_wide.cu on pbs_wide.cuh): 256 emulated threads per
// ciphertext, shared memory as plain arrays, the kernel's two alternating exchange buffers per polynomial, and the
// phases separated exactly where the kernel has its barriers.  Between the two 256-thread barriers of a step the two
// polynomials are NOT synchronised with each other, so the emulation lets one polynomial run all its phases of such a
// segment before the other starts (`order` picks which, and the thread order inside a phase): a hand-over without a
// barrier would make the result depend on the order.  Built by tests/test_pbs_wide_emul.py.
#include "../../fhe_icp_b200/csrc/pbs_wide.cuh"
#include <cmath>
#include <cstring>
#include <vector>
using namespace fhe::nfft;
using namespace fhe::wfft;

namespace {
struct Thread {
    Twiddles tw;
    uint64_t acc_re[8], acc_im[8];
    double re[8], im[8], gre[8], gim[8];
    Monomials mo;
};
int thread_at(int order, int i) { return order == 1 ? 127 - i : (order == 2 ? (i * 37 + 5) & 127 : i); }
}  // namespace

// forward (dir = +1): in = z_j (untwisted, j = 0..1023) -> out = F[k]; inverse (dir = -1): in = F[k] -> out = z_j
extern "C" int emul_wide_fft(const double* in, int dir, int order, double* out) {
    std::vector<cplx> xa(XBUF_ELEMS), xb(XBUF_ELEMS);
    static Thread th[128];
    for (int u = 0; u < 128; ++u) twiddles_init(th[u].tw, u);
    if (dir > 0) {
        for (int i = 0; i < 128; ++i) {
            const int u = thread_at(order, i);
            for (int a = 0; a < 8; ++a) { th[u].re[a] = in[2 * (u + 128 * a)]; th[u].im[a] = in[2 * (u + 128 * a) + 1]; }
            fwd_stage1(th[u].re, th[u].im, th[u].tw, u, xa.data());
        }
        for (int i = 0; i < 128; ++i) { const int u = thread_at(order, i); fwd_stage2(th[u].tw, u, xa.data(), xb.data()); }
        for (int i = 0; i < 128; ++i) {
            const int u = thread_at(order, i);
            fwd_stage3(u, xb.data(), th[u].re, th[u].im);
            for (int kL = 0; kL < 8; ++kL) { out[2 * (u + 128 * kL)] = th[u].re[kL]; out[2 * (u + 128 * kL) + 1] = th[u].im[kL]; }
        }
    } else {
        for (int i = 0; i < 128; ++i) {
            const int u = thread_at(order, i);
            for (int kL = 0; kL < 8; ++kL) { th[u].re[kL] = in[2 * (u + 128 * kL)]; th[u].im[kL] = in[2 * (u + 128 * kL) + 1]; }
            inv_stage3(u, th[u].re, th[u].im, xa.data());
        }
        for (int i = 0; i < 128; ++i) { const int u = thread_at(order, i); inv_stage2(th[u].tw, u, xa.data(), xb.data()); }
        for (int i = 0; i < 128; ++i) {
            const int u = thread_at(order, i);
            inv_stage1(th[u].tw, u, xb.data(), th[u].re, th[u].im);
            for (int a = 0; a < 8; ++a) { out[2 * (u + 128 * a)] = th[u].re[a]; out[2 * (u + 128 * a) + 1] = th[u].im[a]; }
        }
    }
    return 0;
}

static int emul_wide_impl(const double* key_blocks, const double* key_cols, const uint64_t* ct, int n, int beta, const uint64_t* lut,
                          int order, uint64_t* out);

extern "C" int emul_pbs_mb2_wide(const double* key_blocks /* [pairs][32][3][2][1][2][32][2] */, const uint64_t* ct, int n,
                                 int beta, const uint64_t* lut, int order, uint64_t* out /* N + 1 */) {
    return emul_wide_impl(key_blocks, nullptr, ct, n, beta, lut, order, out);
}
// the pointwise stage of pbs_kernel_mb2_pair: key by output column ([pairs][2][32][6][32][2], the layout of
// bsk2_column_split_kernel), S_own / S_oth first, the other polynomial's spectrum last
extern "C" int emul_pbs_mb2_pair(const double* key_cols, const uint64_t* ct, int n, int beta, const uint64_t* lut, int order,
                                 uint64_t* out) {
    return emul_wide_impl(nullptr, key_cols, ct, n, beta, lut, order, out);
}

static int emul_wide_impl(const double* key_blocks, const double* key_cols, const uint64_t* ct, int n, int beta, const uint64_t* lut,
                          int order, uint64_t* out) {
    if (n % 2) return 1;
    std::vector<cplx> omega(128);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int x = 0; x < 64; ++x) {
        omega[x].x = (double)cosl(two_pi * x / 4096.0L);             omega[x].y = (double)sinl(two_pi * x / 4096.0L);
        omega[64 + x].x = (double)cosl(two_pi * (64 * x) / 4096.0L); omega[64 + x].y = (double)sinl(two_pi * (64 * x) / 4096.0L);
    }
    std::vector<int> a_tilde(n + 1);
    for (int i = 0; i <= n; ++i) a_tilde[i] = (int)((((ct[i] >> 51) + 1) >> 1) & 4095);
    static Thread th[2][128];
    const int rot = (4096 - a_tilde[n]) & 4095;
    for (int t = 0; t < 2; ++t) for (int u = 0; u < 128; ++u) {
        twiddles_init(th[t][u].tw, u);
        for (int a = 0; a < 8; ++a) {
            const int j = u + 128 * a;
            uint64_t v0 = 0, v1 = 0;
            if (t == 1) {
                int src = (j - rot) & 4095; v0 = lut[src & 2047]; if (src & 2048) v0 = 0 - v0;
                src = (j + 1024 - rot) & 4095; v1 = lut[src & 2047]; if (src & 2048) v1 = 0 - v1;
            }
            th[t][u].acc_re[a] = v0; th[t][u].acc_im[a] = v1;
        }
    }
    // exchange buffers [t][2] (e0: forward stage 1 / inverse stage 3, e1: forward stage 2 / inverse stage 2) and the spectrum
    // mailbox [t] (tensor memory in the kernel), poisoned so that a read of something never written shows
    std::vector<cplx> mail[2] = {std::vector<cplx>(1024), std::vector<cplx>(1024)};
    for (int t = 0; t < 2; ++t) for (auto& e : mail[t]) e.x = e.y = NAN;
    std::vector<cplx> xbuf[2][2];
    for (int t = 0; t < 2; ++t) for (int q = 0; q < 2; ++q) { xbuf[t][q].resize(XBUF_ELEMS); for (auto& e : xbuf[t][q]) e.x = e.y = NAN; }
    const int pairs = n / 2;
    const int polys[2] = {order == 1 ? 1 : 0, order == 1 ? 0 : 1};
#define FOR_THREADS(t) for (int ii = 0; ii < 128; ++ii) { const int u = thread_at(order, ii); Thread& T = th[t][u]; (void)T;
#define END_THREADS }
    for (int i = 0; i <= pairs; ++i) {
        // ---- segment A: [inverse stage 2 | bar t | inverse stage 1 + accumulate of step i-1] -> stage 1 | bar t | stage 2 | bar t | stage 3 + publish
        for (int pi = 0; pi < 2; ++pi) {
            const int t = polys[pi];
            cplx* e0 = xbuf[t][0].data();
            cplx* e1 = xbuf[t][1].data();
            if (i > 0) {    // inverse stage 2 of step i-1
                FOR_THREADS(t) inv_stage2(T.tw, u, e0, e1); END_THREADS   // bar t
            }
            FOR_THREADS(t)
                if (i > 0) {
                    inv_stage1(T.tw, u, e1, T.re, T.im);
                    for (int a = 0; a < 8; ++a) { T.acc_re[a] += f64_to_torus_u64(T.re[a]); T.acc_im[a] += f64_to_torus_u64(T.im[a]); }
                }
                if (i < pairs) {
                    for (int a = 0; a < 8; ++a) {
                        T.re[a] = top_digit((uint32_t)(T.acc_re[a] >> 32), beta);
                        T.im[a] = top_digit((uint32_t)(T.acc_im[a] >> 32), beta);
                    }
                    fwd_stage1(T.re, T.im, T.tw, u, e0);
                }
            END_THREADS   // bar t
            if (i == pairs) continue;
            FOR_THREADS(t) fwd_stage2(T.tw, u, e0, e1); END_THREADS   // bar t
            FOR_THREADS(t)
                fwd_stage3(u, e1, T.re, T.im);
                for (int kL = 0; kL < 8; ++kL) { cplx v; v.x = T.re[kL]; v.y = T.im[kL]; mail[t][kL * WT + u] = v; }
                monomials_init(T.mo, omega.data(), a_tilde[2 * i], a_tilde[2 * i + 1], u);
            END_THREADS
        }
        if (i == pairs) break;
        // ---- __syncthreads; segment B: pointwise (reads the other polynomial's spectrum) -> inverse stage 3
        const cplx* key_pair = key_blocks ? reinterpret_cast<const cplx*>(key_blocks) + (size_t)i * 32 * MB2_BLOCK_ELEMS : nullptr;
        for (int pi = 0; pi < 2; ++pi) {
            const int t = polys[pi];
            const cplx* o0 = mail[1 - t].data();
            cplx* e0 = xbuf[t][0].data();
            FOR_THREADS(t)
                for (int kL = 0; kL < 8; ++kL) {
                    const cplx fo = o0[kL * WT + u];
                    cplx fa; fa.x = T.re[kL]; fa.y = T.im[kL];
                    if (key_pair) {
                        const cplx* blk = key_pair + (size_t)(4 * kL + (u >> 5)) * MB2_BLOCK_ELEMS;   // slice kL / 2, block of this warp
                        pointwise_bin(t, u & 31, fa, fo, blk, T.mo, T.gre[kL], T.gim[kL]);
                    } else {
                        const cplx* blkc = reinterpret_cast<const cplx*>(key_cols) +
                                           ((size_t)(i * 2 + t) * 32 + (size_t)(4 * kL + (u >> 5))) * COL_BLOCK_ELEMS;
                        cplx so, st;
                        pointwise_sums(t, u & 31, blkc, T.mo, so, st);
                        const double pre = fma(fa.x, so.x, -(fa.y * so.y)), pim = fma(fa.x, so.y, fa.y * so.x);
                        T.gre[kL] = fma(fo.x, st.x, fma(-fo.y, st.y, pre));
                        T.gim[kL] = fma(fo.x, st.y, fma(fo.y, st.x, pim));
                    }
                }
                inv_stage3(u, T.gre, T.gim, e0);
            END_THREADS
        }
        // ---- __syncthreads; the rest of the step (inverse stage 2 | bar t | inverse stage 1 + accumulate) runs into the next
        // step's segment A without another 256-thread barrier, so it is emulated there
    }
#undef FOR_THREADS
#undef END_THREADS
    for (int t = 0; t < 2; ++t) for (int u = 0; u < 128; ++u) for (int a = 0; a < 8; ++a) for (int part = 0; part < 2; ++part) {
        const int x = u + 128 * a + 1024 * part;
        const uint64_t v = part ? th[t][u].acc_im[a] : th[t][u].acc_re[a];
        if (t == 0) { if (x == 0) out[0] = v; else out[NPOLY - x] = 0 - v; }
        else if (x == 0) out[NPOLY] = v;
    }
    return 0;
}
