// tools/ffma2_check.cu -- CPU check that the packed (FFMA2) form of the Gram accumulation (csrc/gram.cuh) produces, bit for bit,
// the sums of the scalar form: same index logic, host emulation of fma.rn.f32x2 (two fmaf).  No GPU needed:
//   nvcc -O2 -std=c++17 -I scalable-bayesian-matrix-factorization_b200/csrc -o /tmp/ffma2_check tools/ffma2_check.cu && /tmp/ffma2_check
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gram.cuh"

using namespace sbmf;

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static float rnd()
{
    s_state ^= s_state << 13;
    s_state ^= s_state >> 7;
    s_state ^= s_state << 17;
    return (float)((double)(s_state >> 11) / 9007199254740992.0 * 4.0 - 2.0);
}

int main()
{
    // every (k, l) of the upper triangle is reachable through exactly the pair the packed form accumulates it in
    bool seen[NACC] = {};
    for (int k = 0; k < 8; ++k)
        for (int l = k; l < 8; ++l) {
            const int p = pi(k, l / 2);
            if (p < 4 || p >= NPAIR) return printf("pair index out of range for (%d,%d)\n", k, l), 1;
            if (seen[gi(k, l)]) return printf("gi collision at (%d,%d)\n", k, l), 1;
            seen[gi(k, l)] = true;
        }
    int bad = 0;
    for (int trial = 0; trial < 200; ++trial) {
        float acc[NACC], out[NACC];
        float2 a2[NPAIR];
        memset(acc, 0, sizeof(acc));
        memset(a2, 0, sizeof(a2));
        const int n = 1 + trial * 7 % 97;
        for (int r = 0; r < n; ++r) {
            f8 f;
            for (int k = 0; k < 8; ++k) f.v[k] = rnd() * (trial % 3 == 0 ? 1e-3f : 1.f);
            const float e = rnd();
            accumulate_scalar(acc, f, e);
            accumulate_packed(a2, f, e);
        }
        unpack_pairs(a2, out);
        for (int i = 0; i < NACC; ++i)
            if (memcmp(&acc[i], &out[i], 4) != 0) {
                if (bad < 5) printf("trial %d entry %d: scalar %.9g packed %.9g\n", trial, i, acc[i], out[i]);
                ++bad;
            }
    }
    // the paired dot products of the REFRESH kernels: (<f, d>, <f, u>) as one chain of pairs == two scalar chains (dot8 of kernels.cu)
    for (int trial = 0; trial < 200; ++trial) {
        f8 f;
        float d[8], u[8];
        float2 du[8];
        for (int k = 0; k < 8; ++k) {
            f.v[k] = rnd();
            d[k] = rnd() * 1e-2f;
            u[k] = rnd();
            du[k] = make_float2(d[k], u[k]);
        }
        float fd = 0.f, fu = 0.f;
        for (int k = 0; k < 8; ++k) fd = fmaf(f.v[k], d[k], fd);
        for (int k = 0; k < 8; ++k) fu = fmaf(f.v[k], u[k], fu);
        const float2 s2 = dot8_pair(f, du);
        if (memcmp(&s2.x, &fd, 4) != 0 || memcmp(&s2.y, &fu, 4) != 0) ++bad;
    }
    if (bad) return printf("FAILED: %d mismatching sums\n", bad), 1;
    printf("ffma2_check ok: packed and scalar Gram accumulation agree bit for bit (200 trials, 44 sums each)\n");
    return 0;
}
