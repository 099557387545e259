// synth_main.cpp -- bin/sbmf_synth: the host-side synthetic rating generator (csrc/synth_host.cpp, compiled into this program;
// NO dependency on libsbmf_cuda.so or CUDA) as a stand-alone tool that writes `user<TAB>item<TAB>rating` triple files the way
// the reference's programs read them (gibbs_sbpmf2.cpp:35-73).  bench.py's reference arm uses it to produce its CPU sample, so
// that the process timing the reference's own program maps none of this repository's GPU code.
//
//   sbmf_synth -users I -items J -ratings N [-test_frac 0.1] [-seed S] [-s_user 0.8] [-s_item 1.0] [-threads T]
//              [-binary 1]      write libFM's binary format (FILE.x + FILE.y, rows `rating user:1 (I+item):1`) instead of triples
//              [-libfm_text 1]  write libFM's text format `rating user:1 (I+item):1` instead of triples
//              [-max_train M]   keep only the first whole users holding about M train ratings (a bounded sample; 0 = all)
//              -train FILE -test FILE
// The last test line pins the id space (num_users = 1 + max user, num_items = J) as the reference sizes its arrays by the
// largest id it reads over train U test ([T]:45-52, 112-119, 152-153).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sbmf_cuda.h"

// `%u\t%u\t%g\n` per rating, formatted by hand (ratings are half stars: "3" or "3.5") on all host threads, written in order
static uint32_t g_libfm_item_offset = 0;   // > 0: write libFM text `rating user:1 (offset + item):1` instead of triples
static inline char* put_uint(char* p, uint32_t v)
{
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

static void write_triples(const char* path, const uint32_t* u, const uint32_t* i, const float* r, uint64_t n, bool pin, uint32_t pin_u, uint32_t pin_i)
{
    FILE* f = fopen(path, "w");
    if (!f) {
        fprintf(stderr, "sbmf_synth: cannot open %s\n", path);
        exit(1);
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const uint64_t T = hw ? hw : 4;
    const uint64_t CH = 1u << 20;                        // lines per chunk
    for (uint64_t c0 = 0; c0 < n; c0 += CH * T) {        // T chunks formatted in parallel, then written in order
        std::vector<std::string> out(T);
        std::vector<std::thread> th;
        for (uint64_t t = 0; t < T; ++t) {
            const uint64_t b = c0 + t * CH, e = std::min(n, b + CH);
            if (b >= e) break;
            th.emplace_back([&, b, e, t]() {
                std::string& o = out[t];
                o.resize((e - b) * 40);
                char* p = &o[0];
                for (uint64_t k = b; k < e; ++k) {
                    auto put_rating = [&]() {
                        const float x2 = r[k] * 2.0f;
                        const int h = (int)x2;
                        if (r[k] >= 0.f && (float)h == x2 && h < 2000000000) {
                            p = put_uint(p, (uint32_t)(h >> 1));
                            if (h & 1) { *p++ = '.'; *p++ = '5'; }
                        } else {
                            p += snprintf(p, 32, "%g", (double)r[k]);
                        }
                    };
                    if (g_libfm_item_offset) {
                        put_rating();
                        *p++ = ' ';
                        p = put_uint(p, u[k]);
                        *p++ = ':'; *p++ = '1'; *p++ = ' ';
                        p = put_uint(p, g_libfm_item_offset + i[k]);
                        *p++ = ':'; *p++ = '1';
                    } else {
                        p = put_uint(p, u[k]);
                        *p++ = '\t';
                        p = put_uint(p, i[k]);
                        *p++ = '\t';
                        put_rating();
                    }
                    *p++ = '\n';
                }
                o.resize((size_t)(p - &o[0]));
            });
        }
        for (auto& x : th) x.join();
        for (uint64_t t = 0; t < T; ++t)
            if (!out[t].empty()) fwrite(out[t].data(), 1, out[t].size(), f);
    }
    if (pin && g_libfm_item_offset) fprintf(f, "3 %u:1 %u:1\n", pin_u, g_libfm_item_offset + pin_i);
    else if (pin) fprintf(f, "%u\t%u\t3\n", pin_u, pin_i);
    fclose(f);
}

// libFM's binary design matrix + target vector (fmatrix.h:34-52, tools/convert.cpp:147-187) of matrix-factorisation rows
// `rating user:1 (I + item):1`: path.x = file_header{2, 4, 2n, n, I + J} + n x {u32 2; {u32 user; f32 1}; {u32 I + item; f32 1}}; path.y = {1, 4, n} + n x f32
static void write_binary(const std::string& path, const uint32_t* u, const uint32_t* i, const float* r, uint64_t n, uint32_t I, uint32_t J)
{
    FILE* fx = fopen((path + ".x").c_str(), "wb");
    FILE* fy = fopen((path + ".y").c_str(), "wb");
    if (!fx || !fy) {
        fprintf(stderr, "sbmf_synth: cannot open %s.x / .y\n", path.c_str());
        exit(1);
    }
    struct { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; } fh = {2u, 4u, 2ull * n, (uint32_t)n, I + J};
    fwrite(&fh, sizeof(fh), 1, fx);
    const uint32_t yh[3] = {1u, 4u, (uint32_t)n};
    fwrite(yh, sizeof(yh), 1, fy);
    fwrite(r, 4, n, fy);
    struct Row { uint32_t size; uint32_t id0; float v0; uint32_t id1; float v1; };
    std::vector<Row> rows(1u << 20);
    for (uint64_t b = 0; b < n; b += rows.size()) {
        const uint64_t cnt = std::min<uint64_t>(rows.size(), n - b);
        for (uint64_t k = 0; k < cnt; ++k) rows[k] = Row{2u, u[b + k], 1.0f, I + i[b + k], 1.0f};
        fwrite(rows.data(), sizeof(Row), cnt, fx);
    }
    fclose(fx);
    fclose(fy);
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
    int threads = 0, binary = 0, libfm_text = 0;
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
        else if (k == "-binary") binary = atoi(v);
        else if (k == "-libfm_text") libfm_text = atoi(v);
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
    if (libfm_text) g_libfm_item_offset = spec.num_users;   // items numbered after ALL users, like scripts/triple_format_to_libfm.pl
    if (binary) {   // FILE.x + FILE.y (the id space is explicit in the header's num_cols; no pin row needed, but kept for equal counts)
        write_binary(train, tu, ti, tr, n, umax + 1, spec.num_items);
        std::vector<uint32_t> u2(su, su + nt), i2(si, si + nt);
        std::vector<float> r2(sr, sr + nt);
        u2.push_back(umax); i2.push_back(spec.num_items - 1); r2.push_back(3.0f);
        write_binary(test, u2.data(), i2.data(), r2.data(), nt + 1, umax + 1, spec.num_items);
    } else {
        write_triples(train.c_str(), tu, ti, tr, n, false, 0, 0);
        write_triples(test.c_str(), su, si, sr, nt, true, umax, spec.num_items - 1);
    }
    printf("{\"n_train\": %llu, \"n_test\": %llu, \"num_users\": %u, \"num_items\": %u, \"n_train_full\": %llu, \"n_test_full\": %llu}\n",
           (unsigned long long)n, (unsigned long long)nt + 1, umax + 1, spec.num_items, (unsigned long long)ntr, (unsigned long long)nte);
    return 0;
}
