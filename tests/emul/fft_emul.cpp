// CPU emulation of one warp running the product's 32x32 negacyclic FFT (fhe_icp_b200/csrc/fft.cuh),
// built by tests/test_fft_emul.py with g++ so the index algebra can be checked without a GPU.
#include "../../fhe_icp_b200/csrc/fft.cuh"
#include <vector>
using namespace fhe::nfft;

extern "C" {
// coef: 2048 doubles -> bins: 1024 complex (re,im interleaved), natural order
void emul_forward(const double* coef, double* bins) {
    std::vector<cplx> tw(TILE_ELEMS), buf(TILE_ELEMS);
    fill_twiddle_table(tw.data());
    static double re[32][32], im[32][32];
    for (int lane = 0; lane < 32; ++lane) {
        for (int j2 = 0; j2 < 32; ++j2) { re[lane][j2] = coef[lane + 32 * j2]; im[lane][j2] = coef[lane + 32 * j2 + 1024]; }
        fwd_phase1(re[lane], im[lane], tw.data(), buf.data(), lane);
    }
    for (int lane = 0; lane < 32; ++lane) {
        fwd_phase2(re[lane], im[lane], buf.data(), lane);
        for (int k1 = 0; k1 < 32; ++k1) {
            bins[2 * (lane + 32 * k1)] = re[lane][brev5(k1)];
            bins[2 * (lane + 32 * k1) + 1] = im[lane][brev5(k1)];
        }
    }
}
void emul_inverse(const double* bins, double* coef) {
    std::vector<cplx> tw(TILE_ELEMS), buf(TILE_ELEMS);
    fill_twiddle_table(tw.data());
    static double re[32][32], im[32][32];
    for (int lane = 0; lane < 32; ++lane) {
        for (int k1 = 0; k1 < 32; ++k1) {
            re[lane][brev5(k1)] = bins[2 * (lane + 32 * k1)];
            im[lane][brev5(k1)] = bins[2 * (lane + 32 * k1) + 1];
        }
        inv_phase1(re[lane], im[lane], tw.data(), buf.data(), lane);
    }
    for (int lane = 0; lane < 32; ++lane) {
        inv_phase2(re[lane], im[lane], buf.data(), lane);
        for (int j2 = 0; j2 < 32; ++j2) { coef[lane + 32 * j2] = re[lane][j2]; coef[lane + 32 * j2 + 1024] = im[lane][j2]; }
    }
}
}
