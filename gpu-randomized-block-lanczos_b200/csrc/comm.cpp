#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <vector>

namespace rbl {
namespace {

// Minimal NCCL ABI (stable since NCCL 2.x): opaque communicator, 128-byte unique id, enum values.
typedef struct { char internal[128]; } nccl_uid_t;
enum { kNcclInt8 = 0, kNcclInt64 = 4, kNcclFloat32 = 7, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };

struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string load_error;
};

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !a.lib; ++i) a.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) {
            a.load_error = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
            return;
        }
#define RBL_SYM(field, name)                                                    \
    *(void**)(&a.field) = dlsym(a.lib, name);                                   \
    if (!a.field) { a.load_error = std::string("missing NCCL symbol ") + name; return; }
        RBL_SYM(GetUniqueId, "ncclGetUniqueId")
        RBL_SYM(CommInitRank, "ncclCommInitRank")
        RBL_SYM(CommDestroy, "ncclCommDestroy")
        RBL_SYM(AllReduce, "ncclAllReduce")
        RBL_SYM(AllGather, "ncclAllGather")
        RBL_SYM(Send, "ncclSend")
        RBL_SYM(Recv, "ncclRecv")
        RBL_SYM(GroupStart, "ncclGroupStart")
        RBL_SYM(GroupEnd, "ncclGroupEnd")
        RBL_SYM(GetErrorString, "ncclGetErrorString")
#undef RBL_SYM
    });
    return a;
}

bool ok(int rc, const char* what, std::string& err) {
    if (rc == 0) return true;
    Api& a = api();
    err = std::string("NCCL ") + what + " failed: " + (a.GetErrorString ? a.GetErrorString(rc) : "?");
    return false;
}

bool ready(std::string& err) {
    Api& a = api();
    if (!a.load_error.empty() || !a.lib) {
        err = a.load_error.empty() ? "NCCL not loaded" : a.load_error;
        return false;
    }
    return true;
}

}  // namespace

bool Comm::unique_id(void* uid128, std::string& err) {
    if (!ready(err)) return false;
    nccl_uid_t id;
    if (!ok(api().GetUniqueId(&id), "GetUniqueId", err)) return false;
    std::memcpy(uid128, &id, sizeof(id));
    return true;
}

namespace {
// Communicators are expensive to create (ncclCommInitRank is a collective of ~1 s); a destroyed handle parks
// its communicator here and a later handle reuses it ONLY when it presents the same 128-byte unique id (and
// device, rank, world): the id names the group, so a different group - or a group re-formed after a peer
// restarted, which must come with a fresh id - never picks up a stale communicator.  Callers that want the
// reuse keep one id per group (bench.py, the single-process group handles in capi.cu).
struct Parked { int device, rank, world; char uid[128]; void* comm; };
std::mutex g_comm_mu;
std::vector<Parked> g_parked;
}  // namespace

bool Comm::init(const void* uid128, int rank_, int world_, std::string& err) {
    rank = rank_;
    world = world_;
    if (world <= 1) return true;
    if (!ready(err)) return false;
    if (!uid128) { err = "null NCCL unique id"; return false; }
    int dev = 0;
    cudaGetDevice(&dev);
    device_ = dev;
    std::memcpy(uid_, uid128, 128);
    failed_ = false;
    {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        for (size_t i = 0; i < g_parked.size(); ++i)
            if (g_parked[i].device == dev && g_parked[i].rank == rank && g_parked[i].world == world &&
                std::memcmp(g_parked[i].uid, uid_, 128) == 0) {
                comm_ = g_parked[i].comm;
                g_parked.erase(g_parked.begin() + i);
                return true;
            }
    }
    nccl_uid_t id;
    std::memcpy(&id, uid128, sizeof(id));
    return check(api().CommInitRank(&comm_, world, id, rank), "CommInitRank", err);
}

bool Comm::check(int rc, const char* what, std::string& err) {
    if (ok(rc, what, err)) return true;
    failed_ = true;
    return false;
}

void Comm::destroy() {
    if (comm_) {
        if (failed_) {
            api().CommDestroy(comm_);   // never re-park a communicator that reported an error
        } else {
            std::lock_guard<std::mutex> lk(g_comm_mu);
            Parked p{device_, rank, world, {0}, comm_};
            std::memcpy(p.uid, uid_, 128);
            g_parked.push_back(p);
        }
    }
    comm_ = nullptr;
}

void Comm::release_cached() {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    for (auto& p : g_parked) api().CommDestroy(p.comm);
    g_parked.clear();
}

bool Comm::allreduce_f64(double* buf, size_t count, cudaStream_t st, std::string& err) {
    if (!active() || count == 0) return true;
    return check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, comm_, st), "AllReduce", err);
}
bool Comm::allreduce_f32(float* buf, size_t count, cudaStream_t st, std::string& err) {
    if (!active() || count == 0) return true;
    return check(api().AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, comm_, st), "AllReduce", err);
}
bool Comm::allgather_i64(const int64_t* send, int64_t* recv, size_t count_per_rank, cudaStream_t st, std::string& err) {
    if (!active()) return true;
    return check(api().AllGather(send, recv, count_per_rank, kNcclInt64, comm_, st), "AllGather", err);
}
bool Comm::group_start(std::string& err) { return !active() || check(api().GroupStart(), "GroupStart", err); }
bool Comm::group_end(std::string& err) { return !active() || check(api().GroupEnd(), "GroupEnd", err); }
bool Comm::send_bytes(const void* buf, size_t bytes, int peer, cudaStream_t st, std::string& err) {
    if (!active() || bytes == 0) return true;
    return check(api().Send(buf, bytes, kNcclInt8, peer, comm_, st), "Send", err);
}
bool Comm::recv_bytes(void* buf, size_t bytes, int peer, cudaStream_t st, std::string& err) {
    if (!active() || bytes == 0) return true;
    return check(api().Recv(buf, bytes, kNcclInt8, peer, comm_, st), "Recv", err);
}

}  // namespace rbl
