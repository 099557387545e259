// comm.h -- NCCL over NVLink for the multi-GPU sweep, loaded at run time (dlopen) so that the single-GPU library has
// no link-time dependency on libnccl.  One communicator per handle (= per rank = per GPU); all calls are enqueued on
// the caller's stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sbmf {

struct Comm {
    void* nccl = nullptr;   // ncclComm_t
    int rank = 0, world = 1;
};

// 0 on success; on failure err holds the reason (library missing, symbol missing, NCCL error string)
int comm_unique_id(uint8_t out[128], std::string& err);
int comm_init(Comm& c, const uint8_t id[128], int rank, int world, std::string& err);
void comm_destroy(Comm& c);
int comm_allreduce_sum_f64(Comm& c, double* buf, size_t count, cudaStream_t st, std::string& err);
// all ranks call with the same (offsets, counts); segment q of buf is broadcast from rank q, in place
int comm_allgatherv_f32(Comm& c, float* buf, const size_t* offsets, const size_t* counts, cudaStream_t st, std::string& err);
int comm_allgatherv_f64(Comm& c, double* buf, const size_t* offsets, const size_t* counts, cudaStream_t st, std::string& err);
// several in-place allgatherv's (same counts/offsets pattern, `nseg` base pointers with stride) fused in one NCCL group
int comm_allgatherv_strided_f32(Comm& c, float* base, size_t stride, int nseg, const size_t* offsets, const size_t* counts, cudaStream_t st,
                                std::string& err);
// NCCL groups nest: everything enqueued between begin and end (including the calls above) becomes ONE grouped launch, so
// independent exchanges share the links instead of running back to back
int comm_group_begin(std::string& err);
int comm_group_end(std::string& err);
int comm_alltoallv_f32(Comm& c, const float* send, const size_t* send_off, const size_t* send_cnt, float* recv, const size_t* recv_off,
                       const size_t* recv_cnt, cudaStream_t st, std::string& err);

}  // namespace sbmf
