// main.cpp -- the host program: the reference's gibbs_sbpmf2 ("[T]" = gibbs_sbpmf2.cpp) with its Gibbs loop running on a
// B200 through the extern "C" sbmf_cuda_* boundary (include/sbmf_cuda.h).  Plain C++; no CUDA, no Python.
//
// Drop-in behaviour:
//   * no arguments: [T]'s hard-coded inputs ../../data/ra.train_sbpmf / ra.test_sbpmf ([T]:32, 98), D = 20 ([T]:224),
//     T = 100 ([T]:322), and [T]'s stdout: "number rows =", "number of user =", "number of items =" ([T]:225-227), one
//     "rmse is <v>" per sweep ([T]:635).
//   * libFM's flag syntax and names (src/util/cmdline.h:29-120, src/libfm/libfm.cpp:86-110): -name value | --name value,
//     ',' or ';' list delimiter, duplicate or unknown flag => "ERROR: ..." like libFM's main (libfm.cpp:636-640).
//       -train F -test F     rating files: `user SEP item SEP rating` triples ([T]:35-73), libFM text `y u:1 i:1` (auto-detected), or
//                            libFM binary F.x + F.y as written by the reference's tools/convert (fmatrix.h:34-52, convert.cpp:147-187;
//                            chosen like Data::load does, Data.h:113-117: when F.x and F.y exist)
//       -dim 'k0,k1,K'       K = latent dimension (third field, libFM's k2); k0/k1 are accepted and must be 1 (biases are part of SBMF)
//       -iter T              sweeps (default 100)        -init_stdev s   (default 0.1, [T]:242)
//       -out F               posterior-mean clamped test predictions, one per line (libfm.cpp:629-634, DVector::save)
//       -rlog F              tab-separated per-sweep log (src/util/rlog.h)
//       -seed n              Philox key (libFM parses -seed and ignores it, libfm.cpp:124; [T] never seeds)
//       -method sbmf|mcmc    sbmf (default): [T]'s stdout.  mcmc: libFM's MCMC front-end outputs instead (fm_learn_mcmc_simultaneous.h:
//                            57-62, 143-147, 244-245): "#Iter=%3d\tTrain=..\tTest=.." per sweep and the file test_rmse_<k0><k1><K>_mcmc
//                            in the CWD with one running-mean test RMSE per line; -task r; -verbosity n; -help
//       -method fm_mcmc|fm_als   libFM's GENERAL FM Gibbs sampler (its -method mcmc / als) on any design matrix: csrc/fm_main.h
//     extensions: -do_sampling 0 (conditional-mean updates; libFM's do_sample=false), -stdev_mode ref|sqrt,
//       -burn_in n, -rebuild_every n, -device n, -item_offset n|auto (libFM text/binary: item feature id - offset = item id),
//       -dump_triples F (write the parsed train triples), -dry_run 1 (parse, print the header lines, stop before touching a GPU),
//       -save_state F / -load_state F (checkpoint after the last sweep / continue a chain: sbmf_cuda_set_state, csrc/checkpoint.cpp),
//       -timing F (stage seconds of the job as JSON: parse / create / build / init / sweeps / out),
//       -dump_xt F (the transposed design matrix of the training data in libFM's binary format, as tools/transpose writes it)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "rating_reader.h"
#include "sbmf_cuda.h"
#include "fm_main.h"

namespace {

struct CmdLine {   // same grammar as src/util/cmdline.h:29-74
    std::map<std::string, std::string> help, value;
    static bool parse_name(std::string& s)
    {
        if (!s.empty() && s[0] == '-') {
            s = (s.size() > 1 && s[1] == '-') ? s.substr(2) : s.substr(1);
            return true;
        }
        return false;
    }
    CmdLine(int argc, char** argv)
    {
        for (int i = 1; i < argc; ++i) {
            std::string s(argv[i]);
            if (!parse_name(s)) throw "cannot parse " + s;
            if (value.count(s)) throw "the parameter " + s + " is already specified";
            if (i + 1 < argc) {
                std::string nx(argv[i + 1]);
                // a value may be a negative number: only treat "-x" as a flag if it is not numeric
                std::string probe = nx;
                const bool is_flag = parse_name(probe) && !(nx.size() > 1 && (isdigit((unsigned char)nx[1]) || nx[1] == '.'));
                if (!is_flag) {
                    value[s] = nx;
                    ++i;
                    continue;
                }
            }
            value[s] = "";
        }
    }
    const std::string& reg(const std::string& p, const std::string& h)
    {
        help[p] = h;
        return p;
    }
    bool has(const std::string& p) const { return value.count(p) != 0; }
    std::string get(const std::string& p, const std::string& def) const { return has(p) ? value.at(p) : def; }
    long geti(const std::string& p, long def) const { return has(p) ? atol(value.at(p).c_str()) : def; }
    double getd(const std::string& p, double def) const { return has(p) ? atof(value.at(p).c_str()) : def; }
    void check() const
    {
        for (const auto& kv : value)
            if (!help.count(kv.first)) throw "the parameter " + kv.first + " does not exist";
    }
    void print_help() const
    {
        for (const auto& kv : help) {
            std::cout << "-" << kv.first;
            for (size_t i = kv.first.size() + 1; i < 16; ++i) std::cout << " ";
            std::cout << kv.second << std::endl;
        }
    }
    static std::vector<std::string> split(const std::string& s, const std::string& delim = ";,")
    {
        std::vector<std::string> out;
        std::string cur;
        for (char c : s) {
            if (delim.find(c) != std::string::npos) {
                out.push_back(cur);
                cur.clear();
            } else {
                cur += c;
            }
        }
        out.push_back(cur);
        return out;
    }
};

struct Ratings {
    std::vector<uint32_t> user, item;
    std::vector<float> rating;
    uint32_t user_max = 0, item_max = 0;
};

// [T]:35-73: sscanf("%u%c%u%c%lf"), a line counts iff all 5 conversions succeed.  libFM text lines ("y u:1 i:1") fail the
// third conversion's separator test (':' then a digit is read as the item id of a triple only if there is no ':' form),
// so the format is detected on the first non-empty line by looking for ':'.
bool file_exists(const std::string& p)
{
    std::ifstream f(p.c_str(), std::ios::binary);
    return f.is_open();
}

// libFM binary design matrix: file_header{u32 id = 2; u32 float_size = 4; u64 num_values; u32 num_rows; u32 num_cols} then per row
// u32 size + size x {u32 id; f32 value} (fmatrix.h:34-52, convert.cpp:147-153, 186-187); targets: {u32 1; u32 4; u32 num_rows} + f32s
// (convert.cpp:159-163, 177).  A matrix-factorisation row has exactly two entries: the user feature and the item feature.
void read_binary(const std::string& path, Ratings& out)
{
    std::ifstream fx((path + ".x").c_str(), std::ios::binary), fy((path + ".y").c_str(), std::ios::binary);
    if (!fx.is_open() || !fy.is_open()) throw "unable to open " + path + ".x / .y";
    struct { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; } fh;
    static_assert(sizeof(fh) == 24, "file_header layout");
    fx.read(reinterpret_cast<char*>(&fh), sizeof(fh));
    if (!fx || fh.id != 2) throw "" + path + ".x: not a libFM binary matrix (file id != 2)";
    if (fh.float_size != 4) throw "" + path + ".x: float size " + std::to_string(fh.float_size) + " is not supported";
    uint32_t yh[3];
    fy.read(reinterpret_cast<char*>(yh), sizeof(yh));
    if (!fy || yh[0] != 1 || yh[1] != 4) throw "" + path + ".y: not a libFM binary target vector";
    if (yh[2] != fh.num_rows) throw "" + path + ": .x and .y disagree on the number of rows";
    if (fh.num_values != 2ull * fh.num_rows) throw "" + path + ".x: every row must hold exactly user:1 item:1";
    out.user.resize(fh.num_rows);
    out.item.resize(fh.num_rows);
    out.rating.resize(fh.num_rows);
    fy.read(reinterpret_cast<char*>(out.rating.data()), (std::streamsize)fh.num_rows * 4);
    if (!fy) throw "" + path + ".y: truncated";
    struct Row { uint32_t size; uint32_t id0; float v0; uint32_t id1; float v1; };
    static_assert(sizeof(Row) == 20, "row layout");
    std::vector<Row> rows(1u << 20);   // 20 MB at a time
    for (uint64_t r0 = 0; r0 < fh.num_rows; r0 += rows.size()) {
        const uint64_t cnt = std::min<uint64_t>(rows.size(), fh.num_rows - r0);
        fx.read(reinterpret_cast<char*>(rows.data()), (std::streamsize)(cnt * sizeof(Row)));
        const uint64_t got = (uint64_t)fx.gcount() / sizeof(Row);
        for (uint64_t i = 0; i < cnt; ++i) {
            if (i >= got || rows[i].size != 2) throw "" + path + ".x: row " + std::to_string(r0 + i) + " does not have exactly two entries";
            out.user[r0 + i] = rows[i].id0;
            out.item[r0 + i] = rows[i].id1;
        }
    }
}

void write_predictions(FILE* o, const std::vector<float>& pred)
{
    const size_t n = pred.size();
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t T = std::max<size_t>(1, std::min<size_t>(hw ? hw : 4, n / 65536 + 1));
    std::vector<std::string> chunk(T);
    std::vector<std::thread> th;
    for (size_t t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            const size_t b = n * t / T, e = n * (t + 1) / T;
            std::string& c = chunk[t];
            c.resize((e - b) * 16);
            char* p = &c[0];
            for (size_t k = b; k < e; ++k) {
                p = std::to_chars(p, p + 15, (double)pred[k], std::chars_format::general, 6).ptr;
                *p++ = '\n';
            }
            c.resize((size_t)(p - &c[0]));
        });
    for (auto& x : th) x.join();
    for (const std::string& c : chunk) fwrite(c.data(), 1, c.size(), o);
}

void read_ratings(const std::string& path, Ratings& out, long item_offset, bool& was_libfm)
{
    if (file_exists(path + ".x") && file_exists(path + ".y")) {   // Data.h:113-117
        read_binary(path, out);
        was_libfm = true;
        if (item_offset > 0)
            for (auto& m : out.item) {
                if ((long)m < item_offset) throw std::string("libFM binary: item feature id below -item_offset in ") + path;
                m -= (uint32_t)item_offset;
            }
        for (size_t n = 0; n < out.user.size(); ++n) {
            if (out.user[n] > out.user_max) out.user_max = out.user[n];
            if (out.item[n] > out.item_max) out.item_max = out.item[n];
        }
        return;
    }
    // text: one read, one chunk per host thread, sscanf semantics (rating_reader.h)
    rating_reader::Parsed parsed;
    bool libfm = false;
    uint64_t bad_line = 0;
    if (!rating_reader::read_text(path, parsed, libfm, bad_line)) throw "unable to open " + path;
    if (bad_line) {
        // [T] would skip the line but still advance its rating index and write out of bounds later
        if (!libfm) throw "malformed rating line " + std::to_string(bad_line) + " in " + path;
        throw "libFM text line " + std::to_string(bad_line) + " in " + path + " is not `y user:1 item:1`";
    }
    out.user.swap(parsed.user);
    out.item.swap(parsed.item);
    out.rating.swap(parsed.rating);
    was_libfm = libfm;
    if (libfm && item_offset > 0)
        for (auto& m : out.item) {
            if ((long)m < item_offset) throw std::string("libFM text: item feature id below -item_offset in ") + path;
            m -= (uint32_t)item_offset;
        }
    for (size_t n = 0; n < out.user.size(); ++n) {
        if (out.user[n] > out.user_max) out.user_max = out.user[n];
        if (out.item[n] > out.item_max) out.item_max = out.item[n];
    }
}

void ck(int rc, sbmf_handle* h, const char* what)
{
    if (rc != SBMF_OK) throw std::string(what) + ": " + sbmf_cuda_last_error(h);
}

}  // namespace

int main(int argc, char** argv)
{
    try {
        CmdLine cmd(argc, argv);
        const std::string p_task = cmd.reg("task", "r=regression (the only task of SBMF)");
        const std::string p_train = cmd.reg("train", "filename for training data; default=../../data/ra.train_sbpmf");
        const std::string p_test = cmd.reg("test", "filename for test data; default=../../data/ra.test_sbpmf");
        const std::string p_out = cmd.reg("out", "filename for output (posterior-mean test predictions)");
        const std::string p_dim = cmd.reg("dim", "'k0,k1,k2': k2=number of latent dimensions; default=1,1,20");
        const std::string p_init = cmd.reg("init_stdev", "stdev for initialization of the factors; default=0.1");
        const std::string p_iter = cmd.reg("iter", "number of Gibbs sweeps; default=100");
        const std::string p_method = cmd.reg("method", "learning method (sbmf | mcmc | fm_mcmc | fm_als); default=sbmf");
        const std::string p_verb = cmd.reg("verbosity", "how much infos to print; default=0");
        const std::string p_rlog = cmd.reg("rlog", "write measurements within iterations to a file; default=''");
        const std::string p_seed = cmd.reg("seed", "integer value, default=1");
        const std::string p_help = cmd.reg("help", "this screen");
        const std::string p_samp = cmd.reg("do_sampling", "0 = conditional-mean updates (no noise); default=1");
        const std::string p_sdm = cmd.reg("stdev_mode", "ref = draw with stdev 1/lambda like the reference, sqrt = sqrt(1/lambda); default=ref");
        const std::string p_hyper = cmd.reg("hyper", "t = hyper-parameter updates of gibbs_sbpmf2.cpp (default); ng_s = Normal-Gamma, no biases, as src/libfm/gibbs_sbpmf2.cpp; ng = the same with its line-412 slip corrected");
        const std::string p_burn = cmd.reg("burn_in", "sweeps before predictions are averaged; default=0");
        const std::string p_reb = cmd.reg("rebuild_every", "rebuild the residual every n sweeps; default=1");
        const std::string p_dev = cmd.reg("device", "CUDA device ordinal; default=0");
        const std::string p_off = cmd.reg("item_offset", "libFM text/binary input: item id = item feature id - offset; default=auto");
        const std::string p_dump = cmd.reg("dump_triples", "write the parsed training triples (user item rating) to this file");
        const std::string p_save = cmd.reg("save_state", "write a checkpoint (factors, biases, hyper-parameters, residual, prediction sums) to this file after the last sweep");
        const std::string p_load = cmd.reg("load_state", "continue the chain from this checkpoint instead of initialising; -iter = number of further sweeps");
        const std::string p_xt = cmd.reg("dump_xt", "write the transposed design matrix of the training data (libFM binary .xt, as tools/transpose writes it) from the device-built layout");
        const std::string p_dry = cmd.reg("dry_run", "1 = parse the inputs, print the header lines and stop (no GPU needed)");
        const std::string p_timing = cmd.reg("timing", "write the wall-clock seconds of the job's stages (parse, create, build, init, sweeps, out) to this file as one JSON object");
        // general FM Gibbs (-method fm_mcmc | fm_als, csrc/fm_main.h): libFM's own flags for -method mcmc / als
        cmd.reg("meta", "fm_mcmc / fm_als: filename with the group id of every attribute, one per line");
        cmd.reg("regular", "fm_mcmc / fm_als: 'r' or 'r0,r1,r2' = prior precision of w0, start value of the w / v group precisions");
        cmd.reg("do_multilevel", "fm_mcmc / fm_als: 0 = fixed hyper-parameters; default=1 (fm_als: 0)");
        cmd.reg("dump_design", "fm_mcmc / fm_als: write the parsed training design matrix (case attribute value target) to this file");
        if (cmd.has(p_help)) {
            cmd.print_help();
            return 0;
        }
        cmd.check();
        const std::string method = cmd.get(p_method, "sbmf");
        if (method == "fm_mcmc" || method == "fm_als") return fm_front::run(cmd, method);   // libFM's general FM Gibbs sampler
        if (method != "sbmf" && method != "mcmc" && method != "MCMC")
            throw "unknown method " + method + " (sbmf | mcmc: the SBMF Gibbs sampler; fm_mcmc | fm_als: libFM's general FM Gibbs sampler)";
        if (cmd.get(p_task, "r") != "r") throw std::string("only -task r is supported");
        uint32_t K = 20;
        if (cmd.has(p_dim)) {
            const std::vector<std::string> d = CmdLine::split(cmd.get(p_dim, ""));
            if (d.size() != 3) throw std::string("-dim needs 'k0,k1,k2'");
            if (atoi(d[0].c_str()) != 1 || atoi(d[1].c_str()) != 1) throw std::string("-dim: k0 and k1 must be 1 (global mean and biases are part of SBMF)");
            K = (uint32_t)atoi(d[2].c_str());
        }
        const uint32_t T = (uint32_t)cmd.geti(p_iter, 100);
        const std::string train_file = cmd.get(p_train, "../../data/ra.train_sbpmf");
        const std::string test_file = cmd.get(p_test, "../../data/ra.test_sbpmf");

        const auto t_job = std::chrono::steady_clock::now();
        auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
        double s_parse = 0.0, s_create = 0.0, s_build = 0.0, s_init = 0.0, s_sweeps = 0.0, s_out = 0.0;
        Ratings tr, te;
        bool libfm_tr = false, libfm_te = false;
        long off = 0;
        const std::string offs = cmd.get(p_off, "auto");
        if (offs != "auto") off = atol(offs.c_str());
        read_ratings(train_file, tr, off, libfm_tr);
        read_ratings(test_file, te, off, libfm_te);
        if ((libfm_tr || libfm_te) && offs == "auto") {
            // libFM MF files number items after the users (scripts/triple_format_to_libfm.pl): offset = 1 + max user id over
            // train U test, applied when every item feature id lies beyond it
            const uint32_t umax = tr.user_max > te.user_max ? tr.user_max : te.user_max;
            uint32_t min_item = UINT32_MAX;
            for (uint32_t m : tr.item) min_item = m < min_item ? m : min_item;
            for (uint32_t m : te.item) min_item = m < min_item ? m : min_item;
            off = (min_item != UINT32_MAX && min_item > umax) ? (long)umax + 1 : 0;
            if (off) {
                for (auto& m : tr.item) m -= (uint32_t)off;
                for (auto& m : te.item) m -= (uint32_t)off;
                if (!tr.item.empty()) tr.item_max -= (uint32_t)off;
                if (!te.item.empty()) te.item_max -= (uint32_t)off;
            }
        }
        s_parse = since(t_job);
        const uint32_t user_max = tr.user_max > te.user_max ? tr.user_max : te.user_max;   // [T]:45-52, 112-119
        const uint32_t item_max = tr.item_max > te.item_max ? tr.item_max : te.item_max;
        const uint32_t num_users = user_max + 1, num_items = item_max + 1;                  // [T]:151-153
        std::cout << "number rows =" << tr.user.size() << "\n";                             // [T]:225-227
        std::cout << "number of user =" << num_users << "\n";
        std::cout << "number of items =" << num_items << "\n";

        if (cmd.has(p_dump)) {
            std::ofstream o(cmd.get(p_dump, "").c_str());
            if (!o.is_open()) throw "unable to open " + cmd.get(p_dump, "");
            for (size_t n = 0; n < tr.user.size(); ++n) o << tr.user[n] << "\t" << tr.item[n] << "\t" << tr.rating[n] << "\n";
        }
        if (cmd.geti(p_dry, 0) != 0) return 0;

        sbmf_config cfg;
        sbmf_cuda_config_default(&cfg);
        cfg.K = K;
        cfg.device = (int32_t)cmd.geti(p_dev, 0);
        cfg.seed = (uint64_t)cmd.geti(p_seed, 1);
        cfg.init_stdev = cmd.getd(p_init, 0.1);
        cfg.burn_in = (uint32_t)cmd.geti(p_burn, 0);
        cfg.rebuild_every = (uint32_t)cmd.geti(p_reb, 1);
        const std::string hyp = cmd.get(p_hyper, "t");
        if (hyp != "t" && hyp != "ng_s" && hyp != "ng") throw std::string("-hyper must be t, ng_s or ng");
        cfg.hyper_mode = hyp == "t" ? SBMF_HYPER_REF_T : (hyp == "ng_s" ? SBMF_HYPER_NG_S : SBMF_HYPER_NG);
        const std::string sdm = cmd.get(p_sdm, "ref");
        if (sdm != "ref" && sdm != "sqrt") throw std::string("-stdev_mode must be ref or sqrt");
        cfg.sample_mode = cmd.geti(p_samp, 1) == 0 ? SBMF_SAMPLE_ZERO_NOISE : (sdm == "sqrt" ? SBMF_SAMPLE_SQRT : SBMF_SAMPLE_REF_VAR_AS_STDEV);

        sbmf_handle* h = NULL;
        auto t_stage = std::chrono::steady_clock::now();
        if (sbmf_cuda_create(&cfg, &h) != SBMF_OK) throw std::string("sbmf_cuda_create: ") + sbmf_cuda_last_error(NULL);
        s_create = since(t_stage);
        t_stage = std::chrono::steady_clock::now();
        ck(sbmf_cuda_set_train(h, tr.user.size(), tr.user.data(), tr.item.data(), tr.rating.data(), num_users, num_items), h, "set_train");
        ck(sbmf_cuda_set_test(h, te.user.size(), te.user.data(), te.item.data(), te.rating.data()), h, "set_test");
        s_build = since(t_stage);
        if (cmd.has(p_xt)) {
            // the device storage build IS the transpose: CSR rows = user features, CSC rows = item features (csrc/xt_writer.cpp).
            // Item features are numbered after the users: at the offset of the libFM input, else at num_users like
            // scripts/triple_format_to_libfm.pl; the feature count is that of the training file alone, like tools/convert.
            const uint32_t xoff = (libfm_tr && off > 0) ? (uint32_t)off : num_users;
            const uint32_t nfeat = tr.item.empty() ? num_users : xoff + tr.item_max + 1;
            const size_t n = tr.user.size();
            std::vector<int64_t> rp((size_t)num_users + 1), cp((size_t)num_items + 1);
            std::vector<uint64_t> rid(n ? n : 1), cid(n ? n : 1);
            ck(sbmf_cuda_get_layout(h, rp.data(), NULL, rid.data(), cp.data(), NULL, cid.data(), NULL), h, "get_layout");
            if (sbmf_cuda_write_libfm_xt(cmd.get(p_xt, "").c_str(), nfeat, num_users, num_items, xoff, n, rp.data(), rid.data(), cp.data(),
                                         cid.data()) != SBMF_OK)
                throw std::string(sbmf_cuda_write_libfm_xt_last_error());
        }
        // state arrays of a checkpoint (sbmf_state members point into these)
        const size_t nI = num_users, nJ = num_items, nN = tr.user.size(), nT = te.user.size();
        std::vector<float> cU, cV, cbi, cbj, cmbi, csbi, cmbj, csbj, cE;
        std::vector<double> csu, cmu, csv, cmv, cps;
        auto bind_state = [&](sbmf_state& st) {
            cU.resize(nI * K); cV.resize((size_t)K * nJ); cbi.resize(nI); cbj.resize(nJ); cmbi.resize(nI); csbi.resize(nI);
            cmbj.resize(nJ); csbj.resize(nJ); cE.resize(nN ? nN : 1); csu.resize(K); cmu.resize(K); csv.resize(K); cmv.resize(K);
            cps.resize(nT ? nT : 1);
            memset(&st, 0, sizeof(st));
            st.U = cU.data(); st.V = cV.data(); st.b_i = cbi.data(); st.b_j = cbj.data(); st.mu_b_i = cmbi.data();
            st.sigma_b_i = csbi.data(); st.mu_b_j = cmbj.data(); st.sigma_b_j = csbj.data(); st.sigma_u = csu.data();
            st.mu_u = cmu.data(); st.sigma_v = csv.data(); st.mu_v = cmv.data(); st.E = cE.data();
        };
        uint32_t first_iter = 0;
        t_stage = std::chrono::steady_clock::now();
        if (cmd.has(p_load)) {
            const std::string path = cmd.get(p_load, "");
            sbmf_checkpoint_dims cd;
            if (sbmf_cuda_checkpoint_read_dims(path.c_str(), &cd) != SBMF_OK) throw std::string(sbmf_cuda_checkpoint_last_error());
            if (cd.num_users != num_users || cd.num_items != num_items || cd.K != K || cd.n_train != nN || cd.n_test != nT ||
                cd.hyper_mode != cfg.hyper_mode)
                throw path + ": checkpoint was written for another problem (users/items/K/ratings/-hyper differ)";
            // the continued chain equals the interrupted one only under the same seed and mode flags
            if (cd.seed != cfg.seed || cd.sample_mode != cfg.sample_mode || cd.burn_in != cfg.burn_in || cd.residual_mode != cfg.residual_mode ||
                cd.rebuild_every != cfg.rebuild_every)
                throw path + ": checkpoint was written by a chain with other flags (seed " + std::to_string(cd.seed) + ", sample mode " +
                    std::to_string(cd.sample_mode) + ", burn-in " + std::to_string(cd.burn_in) + ", residual mode " + std::to_string(cd.residual_mode) +
                    ", rebuild_every " + std::to_string(cd.rebuild_every) + "): pass the same flags to continue it";
            sbmf_checkpoint_dims expect;
            memset(&expect, 0, sizeof(expect));
            expect.num_users = num_users; expect.num_items = num_items; expect.K = K; expect.n_train = nN; expect.n_test = nT;
            sbmf_state st;
            bind_state(st);
            int have_ps = 0;
            if (sbmf_cuda_checkpoint_read(path.c_str(), &expect, &st, cps.data(), &have_ps) != SBMF_OK) throw std::string(sbmf_cuda_checkpoint_last_error());
            ck(sbmf_cuda_set_state(h, &st), h, "set_state");
            if (have_ps) ck(sbmf_cuda_set_pred_sum(h, cps.data()), h, "set_pred_sum");
            first_iter = st.sweeps_done;
        } else {
            ck(sbmf_cuda_init_factors(h, NULL, NULL), h, "init_factors");
        }
        ck(sbmf_cuda_set_timing_enabled(h, 0), h, "set_timing_enabled");
        s_init = since(t_stage);

        std::ofstream rlog;
        if (cmd.has(p_rlog) && !cmd.get(p_rlog, "").empty()) {
            rlog.open(cmd.get(p_rlog, "").c_str());
            if (!rlog.is_open()) throw "unable to open " + cmd.get(p_rlog, "");
            // libFM's field names where the quantity exists here (fm_learn_mcmc_simultaneous.h:231-251), then this sampler's own
            rlog << "rmse\trmse_mcmc_this\trmse_mcmc_all\ttime_learn\talpha\tb_0\n";
        }
        const bool libfm_out = (method != "sbmf");
        std::ofstream file_rmse;
        if (libfm_out) {
            file_rmse.open(("test_rmse_11" + std::to_string(K) + "_mcmc").c_str());   // k0 = k1 = 1
            if (!file_rmse.is_open()) throw std::string("unable to open test_rmse_*_mcmc in the current directory");
        }
        const int verbosity = (int)cmd.geti(p_verb, 0);
        t_stage = std::chrono::steady_clock::now();
        for (uint32_t iter = first_iter; iter < first_iter + T; ++iter) {
            const auto t0 = std::chrono::steady_clock::now();
            double rmse = 0.0, rmse_sweep = 0.0;
            ck(sbmf_cuda_sweep(h, 1), h, "sweep");
            ck(sbmf_cuda_eval(h, &rmse, &rmse_sweep), h, "eval");
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!libfm_out) std::cout << "rmse is " << rmse << std::endl;                  // [T]:635
            if (libfm_out || rlog.is_open() || verbosity > 0) {
                sbmf_state st;
                memset(&st, 0, sizeof(st));
                ck(sbmf_cuda_get_state(h, &st), h, "get_state");
                if (libfm_out) {
                    // Train = RMSE of the training residual this sweep started from (sum e^2 of [T]:358)
                    const double rmse_train = tr.user.empty() ? 0.0 : sqrt(st.sum_e2 / (double)tr.user.size());
                    char buf[16];
                    snprintf(buf, sizeof(buf), "%3u", iter);
                    std::cout << "#Iter=" << buf << "\tTrain=" << rmse_train << "\tTest=" << rmse << std::endl;
                    file_rmse << rmse << "\n" << std::flush;
                }
                if (rlog.is_open())
                    rlog << rmse << "\t" << rmse_sweep << "\t" << rmse << "\t" << dt << "\t" << st.alpha << "\t" << st.b_0 << "\n" << std::flush;
                if (verbosity > 0) std::cout << "alpha=" << st.alpha << "\tb_0=" << st.b_0 << "\ttime=" << dt << std::endl;
            }
        }
        s_sweeps = since(t_stage);
        t_stage = std::chrono::steady_clock::now();
        if (cmd.has(p_out)) {
            std::vector<float> pred(te.user.size());
            if (first_iter + T > cfg.burn_in) ck(sbmf_cuda_get_pred(h, pred.data()), h, "get_pred");
            // the bytes of DVector::save (matrix.h:268-277: `out << value << std::endl`, i.e. %g with 6 significant digits), formatted
            // with std::to_chars on all host threads and written once instead of one flushed stream insertion per line
            FILE* o = fopen(cmd.get(p_out, "").c_str(), "wb");
            if (!o) throw "unable to open " + cmd.get(p_out, "");
            write_predictions(o, pred);
            fclose(o);
        }
        s_out = since(t_stage);
        if (cmd.has(p_timing)) {
            std::ofstream o(cmd.get(p_timing, "").c_str());
            if (!o.is_open()) throw "unable to open " + cmd.get(p_timing, "");
            o << "{\"parse_s\": " << s_parse << ", \"create_s\": " << s_create << ", \"build_s\": " << s_build << ", \"init_s\": " << s_init
              << ", \"sweeps_s\": " << s_sweeps << ", \"out_s\": " << s_out << ", \"total_s\": " << since(t_job) << ", \"sweeps\": " << T
              << ", \"n_train\": " << tr.user.size() << ", \"n_test\": " << te.user.size() << ", \"K\": " << K << "}\n";
        }
        if (cmd.has(p_save)) {
            sbmf_state st;
            bind_state(st);
            ck(sbmf_cuda_get_state(h, &st), h, "get_state");
            ck(sbmf_cuda_get_pred_sum(h, cps.data()), h, "get_pred_sum");
            sbmf_checkpoint_dims cd;
            memset(&cd, 0, sizeof(cd));
            cd.num_users = num_users; cd.num_items = num_items; cd.K = K; cd.hyper_mode = cfg.hyper_mode;
            cd.n_train = nN; cd.n_test = nT;
            cd.seed = cfg.seed; cd.sample_mode = cfg.sample_mode; cd.burn_in = cfg.burn_in; cd.residual_mode = cfg.residual_mode;
            cd.rebuild_every = cfg.rebuild_every;
            if (sbmf_cuda_checkpoint_write(cmd.get(p_save, "").c_str(), &cd, &st, cps.data()) != SBMF_OK)
                throw std::string(sbmf_cuda_checkpoint_last_error());
        }
        sbmf_cuda_destroy(h);
    } catch (std::string& e) {
        std::cerr << std::endl << "ERROR: " << e << std::endl;
        return 1;
    } catch (char const*& e) {
        std::cerr << std::endl << "ERROR: " << e << std::endl;
        return 1;
    }
    return 0;
}
