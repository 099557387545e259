/*
 * libfm_cuda_learner.h -- the reference-side binding of INTEGRATION.md section 4, made real: a libFM learner that runs
 * `-method mcmc` / `-method als` on the B200 through include/sbmf_fm_cuda.h.  TEST INFRASTRUCTURE (tests/test_y_reference_patch.py).
 *
 * `make -C oracle ref` compiles the UNMODIFIED src/libfm/libfm.cpp with
 *     g++ -include oracle/shim_random_libfm.h -include oracle/libfm_cuda_learner.h ... -lsbmf_cuda        -> oracle/_ref/libFM_cuda
 * This header pulls in libFM's own learner headers first (their include guards make libfm.cpp's later #includes no-ops), derives
 * fm_learn_cuda from fm_learn_mcmc_simultaneous and overrides the one virtual that does the work, learn() (fm_learn.h:150,
 * fm_learn_mcmc.h:1154); the closing #define makes libfm.cpp's `new fm_learn_mcmc_simultaneous()` ([L]:413) construct it.
 * Everything else is libFM's: command line, Data::load, meta groups, fm_model::init (the initial w and v come from libFM's own
 * draws and are uploaded), -regular, predict() and -out (they read pred_sum_all, which learn() fills in).
 */
#ifndef SBMF_LIBFM_CUDA_LEARNER_H_
#define SBMF_LIBFM_CUDA_LEARNER_H_
#include <cstdlib>
#include <cstdio>
#include <iostream>
#include <string>
#include <iterator>
#include <algorithm>
#include <iomanip>
#include "../util/util.h"
#include "../util/cmdline.h"
#include "../fm_core/fm_model.h"
#include "src/Data.h"
#include "src/fm_learn.h"
#include "src/fm_learn_mcmc_simultaneous.h"

#include <stdint.h>
#include <vector>
#include "sbmf_fm_cuda.h"

class fm_learn_cuda : public fm_learn_mcmc_simultaneous {
	static void rows_from_transpose(DataSubset& d, std::vector<int64_t>& row_ptr, std::vector<uint32_t>& attr, std::vector<float>& x) {
		// -method mcmc loads only the transposed matrix ([L]:140-149): turn it back into rows (cases ascending inside a column)
		row_ptr.assign((size_t)d.num_cases + 1, 0);
		LargeSparseMatrix<DATA_FLOAT>* t = d.data_t;
		for (t->begin(); !t->end(); t->next()) {
			sparse_row<DATA_FLOAT>& col = t->getRow();
			for (uint k = 0; k < col.size; k++) { row_ptr[col.data[k].id + 1]++; }
		}
		for (uint c = 0; c < d.num_cases; c++) { row_ptr[c + 1] += row_ptr[c]; }
		attr.resize((size_t)row_ptr[d.num_cases]);
		x.resize((size_t)row_ptr[d.num_cases]);
		std::vector<int64_t> fill(row_ptr.begin(), row_ptr.end() - 1);
		for (t->begin(); !t->end(); t->next()) {
			uint a = t->getRowIndex();
			sparse_row<DATA_FLOAT>& col = t->getRow();
			for (uint k = 0; k < col.size; k++) {
				int64_t at = fill[col.data[k].id]++;
				attr[(size_t)at] = a;
				x[(size_t)at] = col.data[k].value;
			}
		}
	}
	static void ck(int rc, sbmf_fm_handle* h, const char* what) {
		if (rc != SBMF_OK) { throw std::string(what) + ": " + sbmf_fm_last_error(h); }
	}
	public:
		virtual void learn(DataSubset& train, DataSubset& test) {
			if (task != TASK_REGRESSION || train.relation.dim > 0) { throw "fm_learn_cuda: regression without relations only"; }
			sbmf_fm_config cfg;
			sbmf_fm_config_default(&cfg);
			cfg.num_attr = fm->num_attribute;
			cfg.num_groups = meta->num_attr_groups;
			cfg.K = (uint32_t)fm->num_factor;
			cfg.k0 = fm->k0; cfg.k1 = fm->k1;
			cfg.do_sample = do_sample; cfg.do_multilevel = do_multilevel;
			cfg.init_stdev = fm->init_stdev;
			cfg.reg0 = fm->reg0; cfg.regw = fm->regw; cfg.regv = fm->regv;
			if (getenv("SBMF_SEED")) { cfg.seed = strtoull(getenv("SBMF_SEED"), NULL, 10); }
			sbmf_fm_handle* h = NULL;
			if (sbmf_fm_create(&cfg, &h) != SBMF_OK) { throw std::string("sbmf_fm_create: ") + sbmf_fm_last_error(NULL); }
			std::vector<uint32_t> group(fm->num_attribute);
			for (uint i = 0; i < fm->num_attribute; i++) { group[i] = meta->attr_group(i); }
			ck(sbmf_fm_set_groups(h, group.data()), h, "set_groups");
			std::vector<int64_t> rp; std::vector<uint32_t> at; std::vector<float> xv, y;
			rows_from_transpose(train, rp, at, xv);
			y.resize(train.num_cases); for (uint c = 0; c < train.num_cases; c++) { y[c] = train.target(c); }
			ck(sbmf_fm_set_train(h, train.num_cases, rp.data(), at.data(), xv.data(), y.data()), h, "set_train");
			rows_from_transpose(test, rp, at, xv);
			y.resize(test.num_cases); for (uint c = 0; c < test.num_cases; c++) { y[c] = test.target(c); }
			ck(sbmf_fm_set_test(h, test.num_cases, rp.data(), at.data(), xv.data(), y.data()), h, "set_test");
			std::vector<float> w0(fm->num_attribute), v0((size_t)fm->num_factor * fm->num_attribute);
			for (uint i = 0; i < fm->num_attribute; i++) { w0[i] = (float)fm->w(i); }
			for (int f = 0; f < fm->num_factor; f++) { for (uint i = 0; i < fm->num_attribute; i++) { v0[(size_t)f * fm->num_attribute + i] = (float)fm->v(f, i); } }
			ck(sbmf_fm_init(h, w0.data(), v0.data()), h, "init");
			std::stringstream tag; tag << fm->k0 << fm->k1 << fm->num_factor;
			std::ofstream file_rmse(("test_rmse_" + tag.str() + "_mcmc").c_str());                         // [GS]:57-62
			for (uint i = 0; i < num_iter; i++) {
				double rmse_train = 0, rmse_test_all = 0;
				ck(sbmf_fm_learn(h, 1), h, "learn");
				ck(sbmf_fm_rmse_history(h, i, 1, &rmse_train, &rmse_test_all), h, "rmse_history");
				std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << rmse_train << "\tTest=" << rmse_test_all << std::endl;   // [GS]:244
				file_rmse << rmse_test_all << "\n";
			}
			// hand the model and the running prediction back to libFM: predict() / -out ([G]:355-379, [L]:629-634) work unchanged
			pred_sum_all.setSize(test.num_cases);
			pred_this.setSize(test.num_cases);
			std::vector<double> ps(test.num_cases ? test.num_cases : 1);
			sbmf_fm_state st; memset(&st, 0, sizeof(st));
			st.w = w0.data(); st.v = v0.data(); st.pred_sum = ps.data();
			ck(sbmf_fm_get_state(h, &st), h, "get_state");
			fm->w0 = st.w0; alpha = st.alpha;
			for (uint i = 0; i < fm->num_attribute; i++) { fm->w(i) = w0[i]; }
			for (int f = 0; f < fm->num_factor; f++) { for (uint i = 0; i < fm->num_attribute; i++) { fm->v(f, i) = v0[(size_t)f * fm->num_attribute + i]; } }
			std::vector<float> last(test.num_cases ? test.num_cases : 1, 0.f);
			if (!do_sample && num_iter > 0 && test.num_cases > 0) { ck(sbmf_fm_predict(h, last.data()), h, "predict"); }   // the last prediction ([G]:362-366)
			for (uint c = 0; c < test.num_cases; c++) { pred_sum_all(c) = ps[c]; pred_this(c) = last[c]; }
			sbmf_fm_destroy(h);
		}
};
#define fm_learn_mcmc_simultaneous fm_learn_cuda
#endif
