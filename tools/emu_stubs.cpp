// tools/emu_stubs.cpp -- TEST INFRASTRUCTURE: what the host build of the SBMF sources (tools/emu_include, launch.h) does not compile:
// NCCL (csrc/comm.cu; the host build runs one "GPU") and the device workload generator (csrc/synth.cu).  Every entry point refuses.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "comm.h"
#include "sbmf_cuda.h"

namespace sbmf {
static int no(std::string& err)
{
    err = "no NCCL in the host build";
    return -1;
}
int comm_unique_id(uint8_t*, std::string& err) { return no(err); }
int comm_init(Comm&, const uint8_t*, int, int, std::string& err) { return no(err); }
void comm_destroy(Comm&) {}
int comm_allreduce_sum_f64(Comm&, double*, size_t, cudaStream_t, std::string& err) { return no(err); }
int comm_allgatherv_f32(Comm&, float*, const size_t*, const size_t*, cudaStream_t, std::string& err) { return no(err); }
int comm_allgatherv_f64(Comm&, double*, const size_t*, const size_t*, cudaStream_t, std::string& err) { return no(err); }
int comm_allgatherv_strided_f32(Comm&, float*, size_t, int, const size_t*, const size_t*, cudaStream_t, std::string& err) { return no(err); }
int comm_group_begin(std::string& err) { return no(err); }
int comm_group_end(std::string& err) { return no(err); }
int comm_alltoallv_f32(Comm&, const float*, const size_t*, const size_t*, float*, const size_t*, const size_t*, cudaStream_t, std::string& err) { return no(err); }
}  // namespace sbmf

extern "C" int sbmf_cuda_synth_generate(const sbmf_synth_spec*, uint64_t*, uint64_t*, uint32_t*, uint32_t*, float*, uint32_t*, uint32_t*, float*)
{
    return SBMF_ERR_UNSUPPORTED;
}
extern "C" const char* sbmf_cuda_synth_last_error(void) { return "no device generator in the host build"; }
// hardware probes (csrc/probe.cu) measure a GPU: nothing to emulate
extern "C" int sbmf_cuda_probe(int, uint64_t, uint64_t, sbmf_probe_result*) { return SBMF_ERR_UNSUPPORTED; }
extern "C" const char* sbmf_cuda_probe_last_error(void) { return "no hardware probes in the host build"; }

// marker: bindings refuse a library that exports this unless the test harness asked for the emulation (sbmf.py load_library, bench.py)
extern "C" int sbmf_simt_host_emulation(void) { return 1; }
