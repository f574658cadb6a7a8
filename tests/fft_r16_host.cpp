// Host emulation of the register-resident Stockham passes of pdspy_b200/csrc/fft_r16.cuh: the same functions the
// kernels call, run for every "thread" of one transform in turn (phases separated where the kernels have barriers).
// Built and driven by tests/test_fft_r16_host.py.
#include "../pdspy_b200/csrc/fft_r16.cuh"
#include <vector>

template <int R>
static void last_pass(const double2 *row, const double2 *tw, int n, int Ns, double2 *out)
{
    const int T = n / 16;
    for (int t = 0; t < T; t++) {
        double2 v[16];
        r16::pass_load<R>(row, n, t, v);
        r16::pass_compute<R>(tw, n, Ns, t, v);
        for (int i = 0; i < 16 / R; i++)
            for (int m = 0; m < R; m++) out[r16::out_index(t + i * T, m, Ns, R)] = v[i * R + m];
    }
}

extern "C" int r16_host_fft(const double *in, int logn, double *out)
{
    const int n = 1 << logn, T = n / 16;
    if (logn < 8 || logn > 12) return 1;
    std::vector<double2> tw(n / 2), row(r16::rowlen(n)), regs((size_t)T * 16);
    for (int q = 0; q < n / 2; q++) tw[q] = make_double2(std::cos(2.0 * M_PI * q / n), std::sin(2.0 * M_PI * q / n));
    int Ns = 1;
    for (int p = 0; p < r16::full_passes(logn); p++) {
        for (int t = 0; t < T; t++) {                            // read phase
            double2 *v = &regs[(size_t)t * 16];
            if (p == 0)
                for (int k = 0; k < 16; k++) v[k] = make_double2(in[2 * (t + k * T)], in[2 * (t + k * T) + 1]);
            else
                r16::pass_load<16>(row.data(), n, t, v);
        }
        for (int t = 0; t < T; t++) {                            // (barrier) compute + write phase
            double2 *v = &regs[(size_t)t * 16];
            r16::pass_compute<16>(tw.data(), n, Ns, t, v);
            r16::pass_store<16>(row.data(), n, Ns, t, v);
        }
        Ns *= 16;
    }
    double2 *o = reinterpret_cast<double2 *>(out);
    switch (r16::last_radix(logn)) {
    case 2: last_pass<2>(row.data(), tw.data(), n, Ns, o); break;
    case 4: last_pass<4>(row.data(), tw.data(), n, Ns, o); break;
    case 8: last_pass<8>(row.data(), tw.data(), n, Ns, o); break;
    default: last_pass<16>(row.data(), tw.data(), n, Ns, o); break;
    }
    return 0;
}
