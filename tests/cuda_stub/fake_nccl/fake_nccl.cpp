// A host-only libnccl.so.2 for tests/test_abi_on_fake_runtime.py: the seven entry points libpmf binds with dlsym, for ranks
// that are THREADS of one process on host memory (the fake CUDA runtime next door).  A communicator is a seat in a group
// keyed by the 128-byte unique id; ncclAllReduce is a rendezvous of the group's ranks that really sums (float32 / float64)
// and writes the sum to every rank's receive buffer; calls made between ncclGroupStart and ncclGroupEnd are queued and run
// at ncclGroupEnd in order, as NCCL does.  Every call is logged (rank, count, datatype, in place or not, inside a group).
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Group {
    int nranks = 0, arrived = 0, generation = 0;
    std::vector<const void*> send; std::vector<void*> recv;
    size_t count = 0; int dtype = 0; bool mismatch = false;
    std::mutex m; std::condition_variable cv;
};
struct Comm { std::shared_ptr<Group> g; int rank; };
struct Op { const void* s; void* r; size_t n; int dt, op; Comm* c; };
struct LogRec { int rank; long long count; int dtype, in_place, grouped, nranks; };
std::mutex g_mu;
std::map<std::string, std::shared_ptr<Group>> g_groups;
std::vector<LogRec> g_log;
int g_mismatches = 0, g_live_comms = 0;
thread_local int t_depth = 0;
thread_local std::vector<Op> t_queue;

int run(const Op& o) {
    Group& g = *o.c->g;
    std::unique_lock<std::mutex> l(g.m);
    const int gen = g.generation;
    if (g.arrived == 0) { g.send.assign(g.nranks, nullptr); g.recv.assign(g.nranks, nullptr); g.count = o.n; g.dtype = o.dt; g.mismatch = false; }
    if (o.n != g.count || o.dt != g.dtype || o.op != 0) g.mismatch = true;          // every rank must make the same call; sum only
    g.send[o.c->rank] = o.s; g.recv[o.c->rank] = o.r;
    if (++g.arrived == g.nranks) {
        if (g.mismatch) { std::lock_guard<std::mutex> k(g_mu); ++g_mismatches; }
        else if (o.dt == 7) {
            std::vector<float> sum(o.n, 0.f);
            for (int r = 0; r < g.nranks; ++r) for (size_t i = 0; i < o.n; ++i) sum[i] += static_cast<const float*>(g.send[r])[i];
            for (int r = 0; r < g.nranks; ++r) std::memcpy(g.recv[r], sum.data(), o.n * 4);
        } else if (o.dt == 8) {
            std::vector<double> sum(o.n, 0.0);
            for (int r = 0; r < g.nranks; ++r) for (size_t i = 0; i < o.n; ++i) sum[i] += static_cast<const double*>(g.send[r])[i];
            for (int r = 0; r < g.nranks; ++r) std::memcpy(g.recv[r], sum.data(), o.n * 8);
        } else { std::lock_guard<std::mutex> k(g_mu); ++g_mismatches; }
        g.arrived = 0; ++g.generation;
        g.cv.notify_all();
    } else {
        g.cv.wait(l, [&] { return g.generation != gen; });
    }
    return 0;
}
}  // namespace

extern "C" {
struct Id128 { char bytes[128]; };
int ncclGetUniqueId(void* out) {
    static int counter = 0;
    std::lock_guard<std::mutex> l(g_mu);
    std::memset(out, 0, 128);
    const int c = ++counter;
    std::memcpy(out, "fake-nccl-id", 12);
    std::memcpy(static_cast<char*>(out) + 16, &c, sizeof c);
    return 0;
}
int ncclCommInitRank(void** comm, int nranks, Id128 id, int rank) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return 4;                          // ncclInvalidArgument
    std::lock_guard<std::mutex> l(g_mu);
    auto& g = g_groups[std::string(id.bytes, 128)];
    if (!g) { g = std::make_shared<Group>(); g->nranks = nranks; }
    if (g->nranks != nranks) return 4;
    *comm = new Comm{g, rank};
    ++g_live_comms;
    return 0;
}
int ncclCommDestroy(void* comm) { if (comm) { delete static_cast<Comm*>(comm); std::lock_guard<std::mutex> l(g_mu); --g_live_comms; } return 0; }
const char* ncclGetErrorString(int e) { return e == 0 ? "no error" : "fake nccl error"; }
int ncclGroupStart() { ++t_depth; return 0; }
int ncclGroupEnd() {
    if (t_depth <= 0) return 5;
    if (--t_depth == 0) { std::vector<Op> q; q.swap(t_queue); for (const Op& o : q) run(o); }
    return 0;
}
int ncclAllReduce(const void* s, void* r, size_t n, int dt, int op, void* comm, void* /*stream*/) {
    if (!comm || !s || !r) return 4;
    Comm* c = static_cast<Comm*>(comm);
    { std::lock_guard<std::mutex> l(g_mu); g_log.push_back({c->rank, (long long)n, dt, s == r, t_depth > 0, c->g->nranks}); }
    Op o{s, r, n, dt, op, c};
    if (t_depth > 0) { t_queue.push_back(o); return 0; }
    return run(o);
}
// ---- doors for the test ------------------------------------------------------------------------------------------------
int fake_nccl_log_count(void) { std::lock_guard<std::mutex> l(g_mu); return (int)g_log.size(); }
int fake_nccl_log(int i, long long* out) {       // rank, count, dtype, in_place, grouped, nranks
    std::lock_guard<std::mutex> l(g_mu);
    if (i < 0 || i >= (int)g_log.size()) return -1;
    const LogRec& x = g_log[i];
    long long v[6] = {x.rank, x.count, x.dtype, x.in_place, x.grouped, x.nranks};
    std::memcpy(out, v, sizeof v);
    return 0;
}
void fake_nccl_state(int* out) { std::lock_guard<std::mutex> l(g_mu); out[0] = g_mismatches; out[1] = g_live_comms; }
}
