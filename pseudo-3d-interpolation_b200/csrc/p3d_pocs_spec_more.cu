// More register-resident plans for common slice shapes (powers of two and 10-smooth lengths); same kernels as
// p3d_pocs_spec.cu, instantiated in their own translation unit so that the library builds in parallel.
// Tile shapes follow the tuned configurations of 256 / 1000 / 2000: about 128 - 256 threads per CTA, column tiles of
// at least 4 adjacent columns (32-byte row segments), as many CTAs per SM as registers and shared memory allow.
#include "p3d_pocs_spec_kernels.cuh"

namespace p3d {

typedef LinePlan<128, 16, 16, 8> LP128;
typedef LinePlan<512, 16, 16, 16, 2> LP512;
typedef LinePlan<1024, 16, 16, 16, 4> LP1024;
typedef LinePlan<2048, 16, 16, 16, 8> LP2048;
typedef LinePlan<400, 20, 20, 20> LP400;
typedef LinePlan<500, 10, 10, 10, 5> LP500;
typedef LinePlan<800, 20, 20, 20, 2> LP800;
typedef LinePlan<1600, 20, 20, 20, 4> LP1600;

bool more_register_cols(SpecKernels& k, int n_iline) {
    switch (n_iline) {
        case 128:  P3D_COLS_BULK(LP128, 16, 5, "spec<128,E16,16x8,C16,cp.async>"); return true;
        case 512:  P3D_COLS_BULK(LP512, 8, 3, "spec<512,E16,16x16x2,C8,cp.async>"); return true;
        case 1024: P3D_COLS_BULK(LP1024, 4, 3, "spec<1024,E16,16x16x4,C4,cp.async>"); return true;
        case 2048: P3D_COLS_BULK(LP2048, 2, 3, "spec<2048,E16,16x16x8,C2,cp.async>"); return true;
        case 400:  P3D_COLS_BULK(LP400, 8, 4, "spec<400,E20,20x20,C8,cp.async>"); return true;
        case 500:  P3D_COLS_BULK(LP500, 4, 5, "spec<500,E10,10x10x5,C4,cp.async>"); return true;
        case 800:  P3D_COLS_BULK(LP800, 4, 4, "spec<800,E20,20x20x2,C4,cp.async>"); return true;
        case 1600: P3D_COLS_BULK(LP1600, 4, 2, "spec<1600,E20,20x20x4,C4,cp.async>"); return true;
        default: return false;
    }
}

bool more_register_rows(SpecKernels& k, int n_xline) {
    switch (n_xline) {
        case 128:  P3D_ROWS(LP128, 16, 4, "spec<128,E16,16x8,RB16>"); return true;
        case 512:  P3D_ROWS(LP512, 4, 5, "spec<512,E16,16x16x2,RB4>"); return true;
        case 1024: P3D_ROWS(LP1024, 2, 5, "spec<1024,E16,16x16x4,RB2>"); return true;
        case 2048: P3D_ROWS(LP2048, 1, 5, "spec<2048,E16,16x16x8,RB1>"); return true;
        case 400:  P3D_ROWS(LP400, 8, 4, "spec<400,E20,20x20,RB8>"); return true;
        case 500:  P3D_ROWS(LP500, 2, 8, "spec<500,E10,10x10x5,RB2>"); return true;
        case 800:  P3D_ROWS(LP800, 2, 6, "spec<800,E20,20x20x2,RB2>"); return true;
        case 1600: P3D_ROWS(LP1600, 1, 6, "spec<1600,E20,20x20x4,RB1>"); return true;
        default: return false;
    }
}

}  // namespace p3d
