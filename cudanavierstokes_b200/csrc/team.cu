// libcudns: z-slabs on several GPUs of one box driven from ONE process (one host thread per solver) -- what the reference does with one
// MPI rank per GPU (src/comm.cpp), without MPI.  cudns_team_create wires N solvers (created by the caller with nranks = N, rank =
// 0..N-1, one device each) together through the public C ABI only (so it serves both precisions):
//   * the slabs' state blocks are mapped into each other (cudns_halo_connect, same-process branch: peer access), after which the stage
//     kernel stores its boundary planes straight into the neighbours' ghost planes over NVLink;
//   * scalar reductions (dt MIN/MAX, bulk / forcing / profile / statistics SUMs: allReduceToMin/Sum, comm.cpp:294-335) and the plane
//     exchange of copyField(0)'s ghost fill (updateHaloFive, comm.cpp:114-134) run through host-side callbacks that meet at a barrier.
// Collective calls (cudns_set_state, cudns_advance, cudns_calc_*, cudns_stats_*, ...) must then be made by all N threads in the same order.
#include "../../include/cudns.h"
#include <cuda_runtime.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

namespace cudns_shared { void set_error(const std::string &msg); }

struct cudns_team;
namespace {
struct Member {
    cudns_handle h = nullptr; cudns_team *team = nullptr; int id = 0, device = 0; cudaStream_t stream = nullptr;
    void *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr; size_t halo_bytes = 0;
    std::vector<double> red;
};
}  // namespace
struct cudns_team {
    int n = 0; std::vector<Member> member;
    std::mutex m; std::condition_variable cv; int waiting = 0; unsigned long gen = 0;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
namespace {
void team_allreduce(void *user, double *dev, int cnt, int op) {
    Member *me = (Member *)user; cudns_team *T = me->team;
    cudaSetDevice(me->device);
    me->red.resize((size_t)cnt);
    cudaStreamSynchronize(me->stream);
    cudaMemcpy(me->red.data(), dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost);
    T->barrier();
    std::vector<double> out((size_t)cnt);
    for (int i = 0; i < cnt; i++) {                                        // rank order: every member computes the same bits
        double a = T->member[0].red[i];
        for (int r = 1; r < T->n; r++) { const double b = T->member[r].red[i]; a = op == 0 ? (b < a ? b : a) : op == 1 ? a + b : (b > a ? b : a); }
        out[i] = a;
    }
    T->barrier();                                                          // everybody has read every red[] before it is reused
    cudaMemcpy(dev, out.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice);
}
void team_exchange(void *user, void *stream) {
    Member *me = (Member *)user; cudns_team *T = me->team;
    cudaSetDevice(me->device);
    cudaStreamSynchronize((cudaStream_t)stream);                           // this slab's send blocks are packed
    T->barrier();
    const Member &lo = T->member[(me->id + T->n - 1) % T->n], &up = T->member[(me->id + 1) % T->n];
    cudaMemcpyPeer(me->recv_hi, me->device, up.send_lo, up.device, me->halo_bytes);
    cudaMemcpyPeer(me->recv_lo, me->device, lo.send_hi, lo.device, me->halo_bytes);
    cudaDeviceSynchronize();
    T->barrier();                                                          // nobody repacks before every neighbour has read
}
}  // namespace

extern "C" {

int cudns_team_create(cudns_handle *solvers, int n, cudns_team_handle *out) {
    if (!solvers || !out || n < 2) { cudns_shared::set_error("cudns_team_create: needs n >= 2 solvers"); return CUDNS_EINVAL; }
    cudns_team *T = new cudns_team();
    T->n = n; T->member.resize(n);
    std::vector<cudns_peer_info> info(n);
    int rc = CUDNS_OK;
    for (int r = 0; r < n && !rc; r++) {
        Member &m = T->member[r];
        m.h = solvers[r]; m.team = T; m.id = r;
        void *st = nullptr;
        if ((rc = cudns_halo_local_info(m.h, &info[r])) || (rc = cudns_get_stream(m.h, &st)) ||
            (rc = cudns_halo_buffers(m.h, &m.send_lo, &m.send_hi, &m.recv_lo, &m.recv_hi, &m.halo_bytes))) break;
        m.device = info[r].device; m.stream = (cudaStream_t)st;
        if ((rc = cudns_set_allreduce(m.h, team_allreduce, &m)) || (rc = cudns_set_exchange(m.h, team_exchange, &m))) break;
    }
    for (int r = 0; r < n && !rc; r++) rc = cudns_halo_connect(solvers[r], &info[(r + n - 1) % n], &info[(r + 1) % n]);
    if (rc) { delete T; return rc; }
    *out = T;
    return CUDNS_OK;
}

int cudns_team_destroy(cudns_team_handle t) { delete t; return CUDNS_OK; }

}  // extern "C"
