// More three-pass mixed-radix plans (p3d_fft_mix.cuh): 100 * {14, 15, 18, 21, 22, 30} with Ra = 10 and
// 256 * {3, 5, 6, 7, 9, 10, 12} with Ra = 16, plus the four-pass LinePlans 2500 and 4096.
#include "p3d_pocs_spec_kernels.cuh"

namespace p3d {

typedef MixPlan3<1400, 10, 14> MP1400;
typedef MixPlan3<1500, 10, 15> MP1500;
typedef MixPlan3<1800, 10, 18> MP1800;
typedef MixPlan3<2100, 10, 21> MP2100;
typedef MixPlan3<2200, 10, 22> MP2200;
typedef MixPlan3<3000, 10, 30> MP3000;
typedef LinePlan<2500, 10, 5, 5, 10, 10> LP2500;
typedef LinePlan<4096, 16, 16, 16, 16> LP4096;

bool mix2_register_cols(SpecKernels& k, int n_iline) {
    switch (n_iline) {
        case 1400: P3D_COLS_BULK(MP1400, 4, 2, "mix<1400,10x14x10,C4,cp.async>"); return true;
        case 1500: P3D_COLS_BULK(MP1500, 4, 2, "mix<1500,10x15x10,C4,cp.async>"); return true;
        case 1800: P3D_COLS_BULK(MP1800, 2, 2, "mix<1800,10x18x10,C2,cp.async>"); return true;
        case 2100: P3D_COLS_BULK(MP2100, 2, 1, "mix<2100,10x21x10,C2,cp.async>"); return true;
        case 2200: P3D_COLS_BULK(MP2200, 2, 1, "mix<2200,10x22x10,C2,cp.async>"); return true;
        case 3000: P3D_COLS_BULK(MP3000, 2, 1, "mix<3000,10x30x10,C2,cp.async>"); return true;
        case 2500: P3D_COLS_BULK(LP2500, 2, 2, "spec<2500,E10,5x5x10x10,C2,cp.async>"); return true;
        case 4096: P3D_COLS_BULK(LP4096, 2, 1, "spec<4096,E16,16x16x16,C2,cp.async>"); return true;
        default: return false;
    }
}

bool mix2_register_rows(SpecKernels& k, int n_xline) {
    switch (n_xline) {
        case 1400: P3D_ROWS(MP1400, 1, 6, "mix<1400,10x14x10,RB1>"); return true;
        case 1500: P3D_ROWS(MP1500, 1, 6, "mix<1500,10x15x10,RB1>"); return true;
        case 1800: P3D_ROWS(MP1800, 1, 5, "mix<1800,10x18x10,RB1>"); return true;
        case 2100: P3D_ROWS(MP2100, 1, 4, "mix<2100,10x21x10,RB1>"); return true;
        case 2200: P3D_ROWS(MP2200, 1, 4, "mix<2200,10x22x10,RB1>"); return true;
        case 3000: P3D_ROWS(MP3000, 1, 2, "mix<3000,10x30x10,RB1>"); return true;
        case 2500: P3D_ROWS(LP2500, 1, 4, "spec<2500,E10,5x5x10x10,RB1>"); return true;
        case 4096: P3D_ROWS(LP4096, 1, 3, "spec<4096,E16,16x16x16,RB1>"); return true;
        default: return false;
    }
}

}  // namespace p3d
