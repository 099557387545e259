// synth_main.cpp -- bin/sbmf_synth: the host-side synthetic rating generator (csrc/synth_host.cpp, compiled into this program;
// NO dependency on libsbmf_cuda.so or CUDA) as a stand-alone tool that writes `user<TAB>item<TAB>rating` triple files the way
// the reference's programs read them (gibbs_sbpmf2.cpp:35-73).  bench.py's reference arm uses it to produce its CPU sample, so
// that the process timing the reference's own program maps none of this repository's GPU code.
//
//   sbmf_synth -users I -items J -ratings N [-test_frac 0.1] [-seed S] [-s_user 0.8] [-s_item 1.0] [-threads T]
//              [-max_train M]   keep only the first whole users holding about M train ratings (a bounded sample; 0 = all)
//              -train FILE -test FILE
// The last test line pins the id space (num_users = 1 + max user, num_items = J) as the reference sizes its arrays by the
// largest id it reads over train U test ([T]:45-52, 112-119, 152-153).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sbmf_cuda.h"

static void write_triples(const char* path, const uint32_t* u, const uint32_t* i, const float* r, uint64_t n, bool pin, uint32_t pin_u, uint32_t pin_i)
{
    FILE* f = fopen(path, "w");
    if (!f) {
        fprintf(stderr, "sbmf_synth: cannot open %s\n", path);
        exit(1);
    }
    std::vector<char> buf(1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (uint64_t k = 0; k < n; ++k) fprintf(f, "%u\t%u\t%g\n", u[k], i[k], (double)r[k]);
    if (pin) fprintf(f, "%u\t%u\t3\n", pin_u, pin_i);
    fclose(f);
}

int main(int argc, char** argv)
{
    sbmf_synth_spec spec;
    memset(&spec, 0, sizeof(spec));
    spec.s_user = 0.8;
    spec.s_item = 1.0;
    spec.test_frac = 0.1;
    spec.seed = 20151001;
    uint64_t max_train = 0;
    int threads = 0;
    std::string train, test;
    for (int a = 1; a + 1 < argc; a += 2) {
        const std::string k = argv[a];
        const char* v = argv[a + 1];
        if (k == "-users") spec.num_users = (uint32_t)strtoul(v, nullptr, 10);
        else if (k == "-items") spec.num_items = (uint32_t)strtoul(v, nullptr, 10);
        else if (k == "-ratings") spec.n_ratings = strtoull(v, nullptr, 10);
        else if (k == "-test_frac") spec.test_frac = atof(v);
        else if (k == "-seed") spec.seed = strtoull(v, nullptr, 10);
        else if (k == "-s_user") spec.s_user = atof(v);
        else if (k == "-s_item") spec.s_item = atof(v);
        else if (k == "-threads") threads = atoi(v);
        else if (k == "-max_train") max_train = strtoull(v, nullptr, 10);
        else if (k == "-train") train = v;
        else if (k == "-test") test = v;
        else {
            fprintf(stderr, "sbmf_synth: unknown flag %s\n", k.c_str());
            return 2;
        }
    }
    if (!spec.num_users || !spec.num_items || !spec.n_ratings || train.empty() || test.empty()) {
        fprintf(stderr, "usage: sbmf_synth -users I -items J -ratings N -train FILE -test FILE [-max_train M] [-seed S] [-test_frac f] [-threads T]\n");
        return 2;
    }
    uint64_t ntr = 0, nte = 0;
    uint32_t *tu = nullptr, *ti = nullptr, *su = nullptr, *si = nullptr;
    float *tr = nullptr, *sr = nullptr;
    if (sbmf_cuda_synth_host_generate(&spec, threads, &ntr, &nte, &tu, &ti, &tr, &su, &si, &sr) != SBMF_OK) {
        fprintf(stderr, "sbmf_synth: %s\n", sbmf_cuda_synth_host_last_error());
        return 1;
    }
    // bounded sample: whole users only (the output is sorted by user)
    uint64_t n = ntr, nt = nte;
    if (max_train && max_train < ntr) {
        n = max_train;
        const uint32_t last = tu[n - 1];
        while (n > 0 && tu[n - 1] == last) --n;      // drop the cut user
        if (n == 0) n = max_train;
    }
    const uint32_t umax = n ? tu[n - 1] : 0;
    nt = 0;
    while (nt < nte && su[nt] <= umax) ++nt;
    write_triples(train.c_str(), tu, ti, tr, n, false, 0, 0);
    write_triples(test.c_str(), su, si, sr, nt, true, umax, spec.num_items - 1);
    printf("{\"n_train\": %llu, \"n_test\": %llu, \"num_users\": %u, \"num_items\": %u, \"n_train_full\": %llu, \"n_test_full\": %llu}\n",
           (unsigned long long)n, (unsigned long long)nt + 1, umax + 1, spec.num_items, (unsigned long long)ntr, (unsigned long long)nte);
    return 0;
}
