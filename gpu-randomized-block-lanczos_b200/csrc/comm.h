// NCCL plumbing for the row-sharded solve (the reference is single-GPU).
// NCCL is loaded with dlopen so that the single-GPU path has no link-time dependency and so that the
// process-wide copy already loaded by torch (same SONAME) is reused when the host is Python.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>

namespace rbl {

class Comm {
public:
    int rank = 0, world = 1;
    bool active() const { return world > 1 && comm_ != nullptr; }
    static bool unique_id(void* uid128, std::string& err);
    bool init(const void* uid128, int rank, int world, std::string& err);
    void destroy();                 // parks the communicator for reuse by the next handle
    static void release_cached();   // really destroys parked communicators
    // in-place sum all-reduce of `count` doubles / floats on `st`
    bool allreduce_f64(double* buf, size_t count, cudaStream_t st, std::string& err);
    bool allreduce_f32(float* buf, size_t count, cudaStream_t st, std::string& err);
    bool allgather_i64(const int64_t* send, int64_t* recv, size_t count_per_rank, cudaStream_t st, std::string& err);
    bool group_start(std::string& err);
    bool group_end(std::string& err);
    bool send_bytes(const void* buf, size_t bytes, int peer, cudaStream_t st, std::string& err);
    bool recv_bytes(void* buf, size_t bytes, int peer, cudaStream_t st, std::string& err);

private:
    bool check(int rc, const char* what, std::string& err);
    void* comm_ = nullptr;
    int device_ = 0;
    char uid_[128] = {0};
    bool failed_ = false;
};

}  // namespace rbl
