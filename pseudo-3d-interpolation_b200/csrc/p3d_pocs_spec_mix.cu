// Three-pass mixed-radix register plans N = Ra * Rb * Ra (p3d_fft_mix.cuh) for lengths whose middle factor does not
// divide the elements per thread: 100 * {3, 6, 7, 9, 11, 12, 13} and 400 * {3, 6}.  Same kernels as p3d_pocs_spec.cu.
#include "p3d_pocs_spec_kernels.cuh"

namespace p3d {

typedef MixPlan3<300, 10, 3> MP300;
typedef MixPlan3<600, 10, 6> MP600;
typedef MixPlan3<700, 10, 7> MP700;
typedef MixPlan3<900, 10, 9> MP900;
typedef MixPlan3<1100, 10, 11> MP1100;
typedef MixPlan3<1300, 10, 13> MP1300;
typedef MixPlan3<1200, 20, 3> MP1200;
typedef MixPlan3<2400, 20, 6> MP2400;

bool mix_register_cols(SpecKernels& k, int n_iline) {
    switch (n_iline) {
        case 300:  P3D_COLS_BULK(MP300, 8, 5, "mix<300,10x3x10,C8,cp.async>"); return true;
        case 600:  P3D_COLS_BULK(MP600, 4, 5, "mix<600,10x6x10,C4,cp.async>"); return true;
        case 700:  P3D_COLS_BULK(MP700, 4, 4, "mix<700,10x7x10,C4,cp.async>"); return true;
        case 900:  P3D_COLS_BULK(MP900, 4, 3, "mix<900,10x9x10,C4,cp.async>"); return true;
        case 1100: P3D_COLS_BULK(MP1100, 4, 2, "mix<1100,10x11x10,C4,cp.async>"); return true;
        case 1300: P3D_COLS_BULK(MP1300, 4, 2, "mix<1300,10x13x10,C4,cp.async>"); return true;
        case 1200: P3D_COLS_BULK(MP1200, 4, 2, "mix<1200,20x3x20,C4,cp.async>"); return true;
        case 2400: P3D_COLS_BULK(MP2400, 2, 2, "mix<2400,20x6x20,C2,cp.async>"); return true;
        default: return false;
    }
}

bool mix_register_rows(SpecKernels& k, int n_xline) {
    switch (n_xline) {
        case 300:  P3D_ROWS(MP300, 4, 8, "mix<300,10x3x10,RB4>"); return true;
        case 600:  P3D_ROWS(MP600, 2, 8, "mix<600,10x6x10,RB2>"); return true;
        case 700:  P3D_ROWS(MP700, 2, 6, "mix<700,10x7x10,RB2>"); return true;
        case 900:  P3D_ROWS(MP900, 2, 5, "mix<900,10x9x10,RB2>"); return true;
        case 1100: P3D_ROWS(MP1100, 1, 8, "mix<1100,10x11x10,RB1>"); return true;
        case 1300: P3D_ROWS(MP1300, 1, 7, "mix<1300,10x13x10,RB1>"); return true;
        case 1200: P3D_ROWS(MP1200, 2, 4, "mix<1200,20x3x20,RB2>"); return true;
        case 2400: P3D_ROWS(MP2400, 1, 4, "mix<2400,20x6x20,RB1>"); return true;
        default: return false;
    }
}

}  // namespace p3d
