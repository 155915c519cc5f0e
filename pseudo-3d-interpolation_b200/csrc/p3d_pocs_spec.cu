// Specialised POCS iteration kernels: compile-time line length, register-resident FFTs
// (p3d_fft_reg.cuh), fused epilogues.  Selected per axis length at plan creation; every
// other shape falls back to the generic kernels.
//
//   cols_iter:  thread (c, j) of a tile of C adjacent columns holds rows j + e*T of column c.
//               load (C*8-byte row segments) -> FFT -> threshold(tau_k) -> IFFT -> store.
//   rows_iter:  thread (j, rr) of a tile of RB rows holds columns j + e*T of row rr.
//               load -> IFFT -> x = alpha*d + (1-alpha*m)*y/(N1 N2) -> sum|x| -> [OUT] -> FFT -> store.
#include "p3d_pocs_spec_kernels.cuh"

namespace p3d {

template <typename F> static std::vector<Cx<F>> spec_twiddle_table_t(const std::vector<int>& radices) {
    std::vector<Cx<F>> t;
    long Ns = 1;
    for (size_t p = 0; p < radices.size(); ++p) {
        const int R = radices[p];
        if (Ns > 1) {
            const size_t base = t.size();
            t.resize(base + (size_t)Ns * R);
            for (long k = 0; k < Ns; ++k)
                for (int r = 0; r < R; ++r) {
                    const double ang = -2.0 * M_PI * (double)((r * k) % (Ns * R)) / (double)(Ns * R);
                    t[base + tw_index(R, (int)Ns, r, (int)k)] = cmake<F>((F)cos(ang), (F)sin(ang));
                }
            if (t.size() & 1) t.push_back(cmake<F>(F(0), F(0)));      // every table starts 16-byte aligned
        }
        Ns *= R;
    }
    if (t.empty()) t.push_back(cmake<F>(F(1), F(0)));
    return t;
}
std::vector<Cx<float>> spec_twiddle_table(const std::vector<int>& radices) { return spec_twiddle_table_t<float>(radices); }
std::vector<Cx<double>> spec_twiddle_table64(const std::vector<int>& radices) { return spec_twiddle_table_t<double>(radices); }

typedef LinePlan<1000, 10, 10, 10, 10> LP1000;
typedef LinePlan<1000, 20, 10, 10, 10> LP1000E20;
typedef LinePlan<2000, 10, 10, 10, 10, 2> LP2000;
typedef LinePlan<2000, 20, 20, 10, 10> LP2000E20;
typedef LinePlan<256, 16, 16, 16> LP256;
typedef LinePlan<200, 20, 10, 20> LP200;
typedef MixPlan3<847, 11, 7> MP847;          // 11 x 7 x 11 (config 3's xline axis)

SpecKernels select_spec_kernels(int n_iline, int n_xline, int variant) {
    SpecKernels k;
    switch (n_iline) {      // column transforms have the length of the iline axis
        case 1000:
            if (variant == 1) P3D_COLS(LP1000E20, 4, 2, "spec<1000,E20,10x10x10,C4,2cta>");
            else if (variant == 2) P3D_COLS_BULK(LP1000E20, 4, 2, "spec<1000,E20,10x10x10,C4,2cta,cp.async>");
            else if (variant == 3) P3D_COLS_BULK(LP1000, 4, 3, "spec<1000,E10,10x10x10,C4,3cta,cp.async>");
            else if (variant == 4) P3D_COLS_BULK(LP1000, 4, 2, "spec<1000,E10,10x10x10,C4,2cta,cp.async>");
            else if (variant == 5) P3D_COLS_BULK(LP1000E20, 8, 1, "spec<1000,E20,10x10x10,C8,1cta,cp.async>");
            else if (variant == 6) P3D_COLS_BULK(LP1000E20, 2, 4, "spec<1000,E20,10x10x10,C2,4cta,cp.async>");
            else if (variant == 7) { P3D_COLS_BULK(LP1000E20, 4, 4, "spec<1000,E20,10x10x10,C4,4cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000E20, 4, 4, true, float, true>; }
            else if (variant == 8) { P3D_COLS_BULK(LP1000E20, 4, 3, "spec<1000,E20,10x10x10,C4,3cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000E20, 4, 3, true, float, true>; }
            else if (variant == 9) { P3D_COLS_BULK(LP1000E20, 8, 2, "spec<1000,E20,10x10x10,C8,2cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000E20, 8, 2, true, float, true>; }
            else              P3D_COLS_BULK(LP1000E20, 4, 3, "spec<1000,E20,10x10x10,C4,3cta,cp.async>");
            break;
        case 2000:
            if (variant == 1) P3D_COLS_BULK(LP2000E20, 4, 1, "spec<2000,E20,20x10x10,C4,1cta,cp.async>");
            else if (variant == 2) P3D_COLS_BULK(LP2000, 4, 1, "spec<2000,E10,10x10x10x2,C4,1cta,cp.async>");
            else if (variant == 3) P3D_COLS(LP2000E20, 4, 1, "spec<2000,E20,20x10x10,C4,1cta>");
            else              P3D_COLS_BULK(LP2000E20, 2, 3, "spec<2000,E20,20x10x10,C2,3cta,cp.async>");
            break;
        case 256:
            if (variant == 1) P3D_COLS(LP256, 16, 3, "spec<256,E16,16x16,C16>");
            else if (variant == 2) P3D_COLS_BULK(LP256, 16, 4, "spec<256,E16,16x16,C16,4cta,cp.async>");
            else if (variant == 3) P3D_COLS_BULK(LP256, 16, 3, "spec<256,E16,16x16,C16,cp.async>");
            else              P3D_COLS_BULK(LP256, 8, 6, "spec<256,E16,16x16,C8,6cta,cp.async>");
            break;
        case 847:  P3D_COLS_BULK(MP847, 4, 3, "mix<847,11x7x11,C4,cp.async>"); break;
        case 200:
            if (variant == 1) P3D_COLS(LP200, 16, 4, "spec<200,E20,10x20,C16>");
            else if (variant == 2) P3D_COLS_BULK(LP200, 16, 5, "spec<200,E20,10x20,C16,5cta,cp.async>");
            else              P3D_COLS_BULK(LP200, 16, 4, "spec<200,E20,10x20,C16,cp.async>");
            break;
        default:
            rader_register_cols(k, n_iline, variant);
            if (!k.cols_iter && !more_register_cols(k, n_iline) && !mix_register_cols(k, n_iline) && !mix2_register_cols(k, n_iline))
                mix3_register_cols(k, n_iline);
            break;
    }
    switch (n_xline) {      // row transforms have the length of the xline axis
        case 1000:
            if (variant == 1) P3D_ROWS(LP1000, 8, 1, "spec<1000,E10,10x10x10,RB8,1cta>");
            else if (variant == 2) { k.rows_iter = launch_rows<LP1000, 4, 2, true>; k.rows_name = "spec<1000,E10,10x10x10,RB4,2cta,l2prefetch>"; k.rows_radices = radices_of<LP1000>(); k.pack_mask = launch_pack<LP1000>; k.rows_T = LP1000::T; }
            else if (variant == 3) P3D_ROWS(LP1000, 4, 2, "spec<1000,E10,10x10x10,RB4,2cta>");
            else if (variant == 5) P3D_ROWS(LP1000, 2, 4, "spec<1000,E10,10x10x10,RB2,4cta>");
            else if (variant == 9) P3D_ROWS(LP1000, 1, 7, "spec<1000,E10,10x10x10,RB1,7cta>");
            else if (variant == 10) P3D_ROWS(LP1000E20, 2, 4, "spec<1000,E20,10x10x10,RB2,4cta>");
            else if (variant == 11) P3D_ROWS(LP1000E20, 1, 8, "spec<1000,E20,10x10x10,RB1,8cta>");
            else              P3D_ROWS(LP1000, 1, 8, "spec<1000,E10,10x10x10,RB1,8cta>");
            break;
        case 2000:
            if (variant == 3) P3D_ROWS(LP2000, 4, 1, "spec<2000,E10,10x10x10x2,RB4>");
            else if (variant == 5) P3D_ROWS(LP2000, 1, 4, "spec<2000,E10,10x10x10x2,RB1,4cta>");
            else if (variant == 6) P3D_ROWS(LP2000, 2, 2, "spec<2000,E10,10x10x10x2,RB2,2cta>");
            else              P3D_ROWS(LP2000E20, 1, 4, "spec<2000,E20,20x10x10,RB1,4cta>");
            break;
        case 256:
            if (variant == 1) P3D_ROWS(LP256, 16, 3, "spec<256,E16,16x16,RB16>");
            else if (variant == 2) P3D_ROWS(LP256, 4, 10, "spec<256,E16,16x16,RB4>");
            else P3D_ROWS(LP256, 8, 5, "spec<256,E16,16x16,RB8>");
            break;
        case 847:
            if (variant == 1) P3D_ROWS(MP847, 2, 4, "mix<847,11x7x11,RB2,4cta>");
            else if (variant == 2) P3D_ROWS(MP847, 4, 3, "mix<847,11x7x11,RB4,3cta>");
            else P3D_ROWS(MP847, 3, 4, "mix<847,11x7x11,RB3,4cta>");
            break;
        case 200:
            if (variant == 1) P3D_ROWS(LP200, 16, 4, "spec<200,E20,10x20,RB16>");
            else if (variant == 2) P3D_ROWS(LP200, 6, 8, "spec<200,E20,10x20,RB6>");
            else P3D_ROWS(LP200, 3, 12, "spec<200,E20,10x20,RB3>");
            break;
        default:
            rader_register_rows(k, n_xline, variant);
            if (!k.rows_iter && !more_register_rows(k, n_xline) && !mix_register_rows(k, n_xline) && !mix2_register_rows(k, n_xline))
                mix3_register_rows(k, n_xline);
            break;
    }
    return k;
}

// ---- float64 state mode: the same iteration kernels instantiated for complex128 -----------------------
// (16-byte elements: E = 10 / 8 keeps the line in <= 64 data registers)
typedef MixPlan3<2000, 10, 20> MP2000;      // 10 x 20 x 10: two exchanges per transform instead of the three of 10 x 10 x 10 x 2
typedef LinePlan<256, 8, 8, 8, 4> LP256E8;
typedef LinePlan<200, 10, 10, 10, 2> LP200E10;

#define P3D_COLS64(LP, C, MINB, NAME) do { k.cols_iter = launch_cols<LP, C, MINB, false, double>; \
                                           k.cols_stats = launch_cols_stats64<LP, C, MINB>; k.cols_C = C; \
                                           k.cols_name = NAME; k.cols_radices = radices_of<LP>(); } while (0)
#define P3D_ROWS64(LP, RB, MINB, NAME) do { k.rows_iter = launch_rows<LP, RB, MINB, false, double>; \
                                            k.rows_iter_io32 = launch_rows<LP, RB, MINB, false, double, true>; \
                                            k.rows_init_io32 = launch_rows_init<LP, RB, MINB, double>; \
                                            k.rows_name = NAME; k.rows_radices = radices_of<LP>(); \
                                            k.pack_mask = launch_pack<LP>; k.rows_T = LP::T; } while (0)

SpecKernels64 select_spec_kernels64(int n_iline, int n_xline, int variant) {
    SpecKernels64 k;
    switch (n_iline) {
        case 1000:
            if (variant == 1) P3D_COLS64(LP1000, 4, 1, "spec64<1000,E10,10x10x10,C4,1cta>");
            else if (variant == 2) P3D_COLS64(LP1000, 2, 3, "spec64<1000,E10,10x10x10,C2,3cta>");
            else if (variant == 3) P3D_COLS64(LP1000, 2, 2, "spec64<1000,E10,10x10x10,C2,2cta>");
            else if (variant == 4) { P3D_COLS64(LP1000, 1, 5, "spec64<1000,E10,10x10x10,C1,5cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 1, 5, true, double>; }
            else if (variant == 5) { P3D_COLS64(LP1000, 1, 4, "spec64<1000,E10,10x10x10,C1,4cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 1, 4, true, double>; }
            else if (variant == 6) { P3D_COLS64(LP1000, 2, 2, "spec64<1000,E10,10x10x10,C2,2cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 2, 2, true, double>; }
            else if (variant == 7) { P3D_COLS64(LP1000, 4, 1, "spec64<1000,E10,10x10x10,C4,1cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 4, 1, true, double>; }
            else if (variant == 8) { P3D_COLS64(LP1000, 4, 2, "spec64<1000,E10,10x10x10,C4,2cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000, 4, 2, true, double, true>; }
            else if (variant == 9) { P3D_COLS64(LP1000, 2, 4, "spec64<1000,E10,10x10x10,C2,4cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000, 2, 4, true, double, true>; }
            else if (variant == 10) { P3D_COLS64(LP1000, 2, 3, "spec64<1000,E10,10x10x10,C2,3cta,cp.async,1buf>"); k.cols_iter = launch_cols<LP1000, 2, 3, true, double, true>; }
            else if (variant == 11) { P3D_COLS64(LP1000, 2, 3, "spec64<1000,E10,10x10x10,C2,3cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 2, 3, true, double>; }
            else { P3D_COLS64(LP1000, 2, 3, "spec64<1000,E10,10x10x10,C2,3cta,cp.async>"); k.cols_iter = launch_cols<LP1000, 2, 3, true, double>; }
            break;
        case 2000:
            if (variant == 1) P3D_COLS64(LP2000, 1, 3, "spec64<2000,E10,10x10x10x2,C1,3cta>");
            else if (variant == 2) P3D_COLS64(LP2000, 2, 1, "spec64<2000,E10,10x10x10x2,C2,1cta>");
            else if (variant == 3) { P3D_COLS64(MP2000, 1, 3, "mix64<2000,10x20x10,C1,3cta,cp.async>"); k.cols_iter = launch_cols<MP2000, 1, 3, true, double>; }
            else if (variant == 4) { P3D_COLS64(MP2000, 2, 1, "mix64<2000,10x20x10,C2,1cta,cp.async>"); k.cols_iter = launch_cols<MP2000, 2, 1, true, double>; }
            else if (variant == 5) { P3D_COLS64(MP2000, 1, 2, "mix64<2000,10x20x10,C1,2cta,cp.async>"); k.cols_iter = launch_cols<MP2000, 1, 2, true, double>; }
            else { P3D_COLS64(MP2000, 2, 2, "mix64<2000,10x20x10,C2,2cta,cp.async,1buf>"); k.cols_iter = launch_cols<MP2000, 2, 2, true, double, true>; }
            break;
        case 256:  P3D_COLS64(LP256E8, 8, 3, "spec64<256,E8,8x8x4,C8>"); break;
        case 200:  P3D_COLS64(LP200E10, 8, 4, "spec64<200,E10,10x10x2,C8>"); break;
        case 847:  P3D_COLS64(MP847, 2, 3, "mix64<847,11x7x11,C2,3cta>"); break;
        default: rader_register_cols64(k, n_iline, variant); break;
    }
    switch (n_xline) {
        case 1000:
            if (variant == 1) P3D_ROWS64(LP1000, 2, 2, "spec64<1000,E10,10x10x10,RB2,2cta>");
            else if (variant == 4 || variant == 6) P3D_ROWS64(LP1000, 1, 5, "spec64<1000,E10,10x10x10,RB1,5cta>");
            else if (variant == 5 || variant == 7) P3D_ROWS64(LP1000, 1, 6, "spec64<1000,E10,10x10x10,RB1,6cta>");
            else if (variant == 8) P3D_ROWS64(LP1000, 2, 3, "spec64<1000,E10,10x10x10,RB2,3cta>");
            else if (variant == 9) P3D_ROWS64(LP1000, 3, 2, "spec64<1000,E10,10x10x10,RB3,2cta>");
            else if (variant == 10) P3D_ROWS64(LP1000, 4, 1, "spec64<1000,E10,10x10x10,RB4,1cta>");
            else if (variant == 11) { P3D_ROWS64(LP1000, 1, 4, "spec64<1000,E10,10x10x10,RB1,4cta,l2prefetch>"); k.rows_iter_io32 = launch_rows<LP1000, 1, 4, true, double, true>; }
            else if (variant == 12) P3D_ROWS64(LP1000, 1, 4, "spec64<1000,E10,10x10x10,RB1,4cta>");
            else { P3D_ROWS64(LP1000, 1, 4, "spec64<1000,E10,10x10x10,RB1,4cta,l2prefetch>"); k.rows_iter_io32 = launch_rows<LP1000, 1, 4, true, double, true>; }
            break;
        case 2000:
            if (variant == 1 || variant == 2) P3D_ROWS64(LP2000, 1, 2, "spec64<2000,E10,10x10x10x2,RB1,2cta>");
            else if (variant == 3) P3D_ROWS64(MP2000, 1, 3, "mix64<2000,10x20x10,RB1,3cta>");
            else if (variant == 4) P3D_ROWS64(MP2000, 1, 4, "mix64<2000,10x20x10,RB1,4cta>");
            else P3D_ROWS64(MP2000, 1, 2, "mix64<2000,10x20x10,RB1,2cta>");
            break;
        case 256:  P3D_ROWS64(LP256E8, 4, 5, "spec64<256,E8,8x8x4,RB4>"); break;
        case 200:  P3D_ROWS64(LP200E10, 4, 6, "spec64<200,E10,10x10x2,RB4>"); break;
        case 847:  P3D_ROWS64(MP847, 1, 5, "mix64<847,11x7x11,RB1,5cta>"); break;
        default: break;
    }
    return k;
}

}  // namespace p3d
