// fm_main.h -- host front-end of the general FM Gibbs sampler (`sbmf -method fm_mcmc | fm_als`): libFM's own command line for
// `-method mcmc` / `-method als` ("[L]" = src/libfm/libfm.cpp) over the C ABI of include/sbmf_fm_cuda.h.  Plain C++.
//
//   -train F -test F      design matrices: libFM text `y id:value id:value ...` (Data.h:184-278: blank lines and lines starting
//                         with '#' are skipped) or libFM binary F.x + F.y (fmatrix.h:34-52; chosen when both exist, Data.h:113-117)
//   -dim 'k0,k1,K'        use w0 / use w / number of factors ([L]:377-383)          -iter n   ([L]:415, default 100)
//   -meta F               group id of every attribute, one per line (Data.h:49-61)   -init_stdev s (default 0.1, [L]:127)
//   -regular 'r' | 'r0,r1,r2'   ([L]:485-505)                                        -seed n (Philox key; [L]:124 ignores it)
//   -out F  -rlog F  -verbosity n  -task r
//   fm_als = libFM's `-method als`: do_sampling = do_multilevel = 0 ([L]:131-136); -do_sampling / -do_multilevel override either.
// Outputs as libFM: "#Iter=%3d\tTrain=..\tTest=.." per iteration and test_rmse_<k0><k1><K>_mcmc in the CWD ([GS]:57-62, 244-245),
// -out = fm_learn_mcmc::predict ([G]:355-379) one value per line ([L]:629-634).
// The attribute count is libFM's: max(train.num_feature, test.num_feature) + 1 ([L]:326; a file's num_feature is its largest
// id + 1), i.e. one data-free attribute beyond the largest id, which is drawn from its prior like in libFM.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "sbmf_fm_cuda.h"

namespace fm_front {

struct Design {
    std::vector<int64_t> row_ptr{0};
    std::vector<uint32_t> attr;
    std::vector<float> x, y;
    uint32_t num_feature = 0;   // largest id + 1 (Data.h:205-221), 0 without any feature
};

inline bool exists(const std::string& p)
{
    std::ifstream f(p.c_str(), std::ios::binary);
    return f.is_open();
}

// Data.h:184-278
inline void read_text(const std::string& path, Design& d)
{
    std::ifstream f(path.c_str());
    if (!f.is_open()) throw "unable to open " + path;
    std::string line;
    while (std::getline(f, line)) {
        const char* p = line.c_str();
        while (*p == ' ' || *p == 9) p++;
        if (*p == 0 || *p == '#') continue;
        float v;
        int nchar = 0, feat = 0;
        if (sscanf(p, "%f%n", &v, &nchar) < 1) throw "cannot parse line \"" + line + "\" at character " + p[0];
        p += nchar;
        d.y.push_back(v);
        while (sscanf(p, "%d:%f%n", &feat, &v, &nchar) >= 2) {
            p += nchar;
            if (feat < 0) throw "negative attribute id in line \"" + line + "\"";
            d.attr.push_back((uint32_t)feat);
            d.x.push_back(v);
            if ((uint32_t)feat + 1 > d.num_feature) d.num_feature = (uint32_t)feat + 1;
        }
        while (*p == ' ' || *p == 9 || *p == '\r') p++;
        if (*p != 0 && *p != '#') throw "cannot parse line \"" + line + "\" at character " + p[0];
        d.row_ptr.push_back((int64_t)d.attr.size());
    }
}

// fmatrix.h:34-52 + convert.cpp:147-187
inline void read_binary(const std::string& path, Design& d)
{
    std::ifstream fx((path + ".x").c_str(), std::ios::binary), fy((path + ".y").c_str(), std::ios::binary);
    if (!fx.is_open() || !fy.is_open()) throw "unable to open " + path + ".x / .y";
    struct { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; } fh;
    static_assert(sizeof(fh) == 24, "file_header layout");
    fx.read(reinterpret_cast<char*>(&fh), sizeof(fh));
    if (!fx || fh.id != 2 || fh.float_size != 4) throw "" + path + ".x: not a libFM binary matrix with 4-byte values";
    uint32_t yh[3];
    fy.read(reinterpret_cast<char*>(yh), sizeof(yh));
    if (!fy || yh[0] != 1 || yh[1] != 4 || yh[2] != fh.num_rows) throw "" + path + ".y: not the target vector of " + path + ".x";
    d.y.resize(fh.num_rows);
    fy.read(reinterpret_cast<char*>(d.y.data()), (std::streamsize)fh.num_rows * 4);
    if (!fy) throw "" + path + ".y: truncated";
    d.attr.reserve(fh.num_values);
    d.x.reserve(fh.num_values);
    struct Entry { uint32_t id; float v; };
    std::vector<Entry> row;
    for (uint32_t r = 0; r < fh.num_rows; ++r) {
        uint32_t sz = 0;
        fx.read(reinterpret_cast<char*>(&sz), 4);
        row.resize(sz);
        if (sz) fx.read(reinterpret_cast<char*>(row.data()), (std::streamsize)sz * 8);
        if (!fx) throw "" + path + ".x: truncated at row " + std::to_string(r);
        for (const Entry& en : row) {
            d.attr.push_back(en.id);
            d.x.push_back(en.v);
            if (en.id + 1 > d.num_feature) d.num_feature = en.id + 1;
        }
        d.row_ptr.push_back((int64_t)d.attr.size());
    }
    if (d.attr.size() != fh.num_values) throw "" + path + ".x: header and rows disagree on the number of values";
}

inline void read_design(const std::string& path, Design& d)
{
    if (exists(path + ".x") && exists(path + ".y")) read_binary(path, d);
    else read_text(path, d);
}

inline void ckfm(int rc, sbmf_fm_handle* h, const char* what)
{
    if (rc != SBMF_OK) throw std::string(what) + ": " + sbmf_fm_last_error(h);
}

// CmdLineT: the CmdLine of main.cpp (same grammar as src/util/cmdline.h)
template <class CmdLineT>
int run(const CmdLineT& cmd, const std::string& method)
{
    const bool als = method == "fm_als";
    if (!cmd.has("train") || !cmd.has("test")) throw std::string("-method " + method + " needs -train and -test");
    Design tr, te;
    read_design(cmd.get("train", ""), tr);
    read_design(cmd.get("test", ""), te);
    const uint32_t p = std::max(tr.num_feature, te.num_feature) + 1;   // [L]:326
    sbmf_fm_config cfg;
    sbmf_fm_config_default(&cfg);
    cfg.num_attr = p;
    if (cmd.has("dim")) {
        const std::vector<std::string> d = CmdLineT::split(cmd.get("dim", ""));
        if (d.size() != 3) throw std::string("-dim needs 'k0,k1,k2'");
        cfg.k0 = atoi(d[0].c_str()) != 0;
        cfg.k1 = atoi(d[1].c_str()) != 0;
        cfg.K = (uint32_t)atoi(d[2].c_str());
    }
    std::vector<uint32_t> group(p, 0);
    if (cmd.has("meta")) {
        std::ifstream f(cmd.get("meta", "").c_str());
        if (!f.is_open()) throw "unable to open " + cmd.get("meta", "");
        uint32_t g, i = 0, ng = 0;
        while (f >> g) {
            if (i < p) group[i] = g;
            ++i;
            if (g + 1 > ng) ng = g + 1;
        }
        if (i < tr.num_feature) throw std::string("-meta lists fewer attributes than the training file uses");
        cfg.num_groups = ng ? ng : 1;
    }
    cfg.do_sample = (int32_t)cmd.geti("do_sampling", als ? 0 : 1);
    cfg.do_multilevel = (int32_t)cmd.geti("do_multilevel", als ? 0 : 1);
    cfg.device = (int32_t)cmd.geti("device", 0);
    cfg.seed = (uint64_t)cmd.geti("seed", 1);
    cfg.init_stdev = cmd.getd("init_stdev", 0.1);
    if (cmd.has("regular")) {
        const std::vector<std::string> r = CmdLineT::split(cmd.get("regular", ""));
        if (r.size() == 1) cfg.reg0 = cfg.regw = cfg.regv = atof(r[0].c_str());
        else if (r.size() == 3) { cfg.reg0 = atof(r[0].c_str()); cfg.regw = atof(r[1].c_str()); cfg.regv = atof(r[2].c_str()); }
        else throw std::string("-regular needs 'r' or 'r0,r1,r2' (per-group lists are not supported)");
    }
    const uint32_t T = (uint32_t)cmd.geti("iter", 100);
    std::cout << "#cases train=" << tr.y.size() << "\ttest=" << te.y.size() << "\t#attr=" << p << "\t#groups=" << cfg.num_groups << std::endl;
    if (cmd.has("dump_design")) {   // the parsed training matrix, one `case attr value` per entry (reader tests)
        std::ofstream o(cmd.get("dump_design", "").c_str());
        if (!o.is_open()) throw "unable to open " + cmd.get("dump_design", "");
        o.precision(9);
        for (size_t r = 0; r + 1 < tr.row_ptr.size(); ++r)
            for (int64_t k = tr.row_ptr[r]; k < tr.row_ptr[r + 1]; ++k) o << r << " " << tr.attr[k] << " " << tr.x[k] << " " << tr.y[r] << "\n";
    }
    if (cmd.geti("dry_run", 0) != 0) return 0;
    if (tr.y.size() >= (1ull << 32) || te.y.size() >= (1ull << 32)) throw std::string("more than 2^32-1 cases are not supported");

    sbmf_fm_handle* h = NULL;
    if (sbmf_fm_create(&cfg, &h) != SBMF_OK) throw std::string("sbmf_fm_create: ") + sbmf_fm_last_error(NULL);
    if (cmd.has("meta")) ckfm(sbmf_fm_set_groups(h, group.data()), h, "set_groups");
    ckfm(sbmf_fm_set_train(h, (uint32_t)tr.y.size(), tr.row_ptr.data(), tr.attr.data(), tr.x.data(), tr.y.data()), h, "set_train");
    ckfm(sbmf_fm_set_test(h, (uint32_t)te.y.size(), te.row_ptr.data(), te.attr.data(), te.x.data(), te.y.data()), h, "set_test");
    ckfm(sbmf_fm_init(h, NULL, NULL), h, "init");
    const std::string tag = std::to_string(cfg.k0) + std::to_string(cfg.k1) + std::to_string(cfg.K);
    std::ofstream file_rmse(("test_rmse_" + tag + "_mcmc").c_str());   // [GS]:57-62
    if (!file_rmse.is_open()) throw std::string("unable to open test_rmse_*_mcmc in the current directory");
    std::ofstream rlog;
    if (cmd.has("rlog") && !cmd.get("rlog", "").empty()) {
        rlog.open(cmd.get("rlog", "").c_str());
        if (!rlog.is_open()) throw "unable to open " + cmd.get("rlog", "");
        rlog << "rmse\trmse_mcmc_all\trmse_train\talpha\tw0\n";      // libFM's field names where the quantity exists ([GS]:246-251)
    }
    for (uint32_t it = 0; it < T; ++it) {
        double a = 0.0, b = 0.0;
        ckfm(sbmf_fm_learn(h, 1), h, "learn");
        ckfm(sbmf_fm_rmse_history(h, it, 1, &a, &b), h, "rmse_history");
        char buf[16];
        snprintf(buf, sizeof(buf), "%3u", it);
        std::cout << "#Iter=" << buf << "\tTrain=" << a << "\tTest=" << b << std::endl;   // [GS]:244
        file_rmse << b << "\n" << std::flush;
        if (rlog.is_open()) {
            sbmf_fm_state st;
            memset(&st, 0, sizeof(st));
            ckfm(sbmf_fm_get_state(h, &st), h, "get_state");
            rlog << b << "\t" << b << "\t" << a << "\t" << st.alpha << "\t" << st.w0 << "\n" << std::flush;
        }
    }
    if (cmd.has("out") && T > 0 && !te.y.empty()) {
        std::vector<float> pred(te.y.size());
        ckfm(sbmf_fm_predict(h, pred.data()), h, "predict");
        std::ofstream o(cmd.get("out", "").c_str());
        if (!o.is_open()) throw "unable to open " + cmd.get("out", "");
        for (float v : pred) o << (double)v << "\n";
    }
    sbmf_fm_destroy(h);
    return 0;
}

}  // namespace fm_front
