// Three-pass mixed-radix plans with Ra = 16 (p3d_fft_mix.cuh): 256 * {3, 5, 6, 7, 9, 10, 12}.
#include "p3d_pocs_spec_kernels.cuh"

namespace p3d {

typedef MixPlan3<768, 16, 3> MP768;
typedef MixPlan3<1280, 16, 5> MP1280;
typedef MixPlan3<1536, 16, 6> MP1536;
typedef MixPlan3<1792, 16, 7> MP1792;
typedef MixPlan3<2304, 16, 9> MP2304;
typedef MixPlan3<2560, 16, 10> MP2560;
typedef MixPlan3<3072, 16, 12> MP3072;

bool mix3_register_cols(SpecKernels& k, int n_iline) {
    switch (n_iline) {
        case 768:  P3D_COLS_BULK(MP768, 4, 4, "mix<768,16x3x16,C4,cp.async>"); return true;
        case 1280: P3D_COLS_BULK(MP1280, 4, 2, "mix<1280,16x5x16,C4,cp.async>"); return true;
        case 1536: P3D_COLS_BULK(MP1536, 4, 2, "mix<1536,16x6x16,C4,cp.async>"); return true;
        case 1792: P3D_COLS_BULK(MP1792, 2, 3, "mix<1792,16x7x16,C2,cp.async>"); return true;
        case 2304: P3D_COLS_BULK(MP2304, 2, 2, "mix<2304,16x9x16,C2,cp.async>"); return true;
        case 2560: P3D_COLS_BULK(MP2560, 2, 2, "mix<2560,16x10x16,C2,cp.async>"); return true;
        case 3072: P3D_COLS_BULK(MP3072, 2, 2, "mix<3072,16x12x16,C2,cp.async>"); return true;
        default: return false;
    }
}

bool mix3_register_rows(SpecKernels& k, int n_xline) {
    switch (n_xline) {
        case 768:  P3D_ROWS(MP768, 2, 6, "mix<768,16x3x16,RB2>"); return true;
        case 1280: P3D_ROWS(MP1280, 1, 6, "mix<1280,16x5x16,RB1>"); return true;
        case 1536: P3D_ROWS(MP1536, 1, 5, "mix<1536,16x6x16,RB1>"); return true;
        case 1792: P3D_ROWS(MP1792, 1, 4, "mix<1792,16x7x16,RB1>"); return true;
        case 2304: P3D_ROWS(MP2304, 1, 4, "mix<2304,16x9x16,RB1>"); return true;
        case 2560: P3D_ROWS(MP2560, 1, 3, "mix<2560,16x10x16,RB1>"); return true;
        case 3072: P3D_ROWS(MP3072, 1, 3, "mix<3072,16x12x16,RB1>"); return true;
        default: return false;
    }
}

}  // namespace p3d
