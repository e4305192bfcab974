// Analysis tool (not product, not test): per-row dopri5 attempt / rejection counts of the dynamics
// stage for a batch of parameter sets read from a raw vag_params file.
//   g++ -std=c++17 -O2 -DVAG_INSTRUMENT -x c++ scripts/step_stats.cpp -o /tmp/step_stats
//   /tmp/step_stats params.bin t_min t_max
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../oracle/hostemu/hostemu.cpp"

int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END);
    const size_t n = ftell(f) / sizeof(vag_params);
    fseek(f, 0, SEEK_SET);
    std::vector<vag_params> P(n);
    if (fread(P.data(), sizeof(vag_params), n, f) != n) return 1;
    const double t_min = atof(argv[2]), t_max = atof(argv[3]);
    for (size_t i = 0; i < n; ++i) {
        HostBatch hb;
        g_step_stats = StepStats{};
        run_front(hb, &P[i], 1, t_min, t_max);
        printf("%zu %d %d %ld %ld %ld %ld\n", i, hb.w.hdr[0].n_t, hb.w.totals[TOT_ROWS], g_step_stats.attempts, g_step_stats.rejects,
               g_step_stats.attempts_s, g_step_stats.rejects_s);
    }
}
