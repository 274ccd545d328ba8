// pqp_multi.cu -- multi-GPU entry points of libpqp.so (include/pqp_multi.h): shards over the GPUs of one box, one NCCL
// all-gather of the solved Frenet states at the end.  Host orchestration only.  NCCL is resolved with dlopen / dlsym
// so that the library carries no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/pqp_multi.h"
#include "pqp_handle.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

const NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the copy already in the process (PyTorch's) first: two NCCL instances in one process would not share state
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return;
        api.lib = lib;
#define PQP_SYM(field, name) api.field = (decltype(api.field))dlsym(lib, name)
        PQP_SYM(GetUniqueId, "ncclGetUniqueId");
        PQP_SYM(CommInitRank, "ncclCommInitRank");
        PQP_SYM(CommInitAll, "ncclCommInitAll");
        PQP_SYM(CommDestroy, "ncclCommDestroy");
        PQP_SYM(AllGather, "ncclAllGather");
        PQP_SYM(GroupStart, "ncclGroupStart");
        PQP_SYM(GroupEnd, "ncclGroupEnd");
        PQP_SYM(GetErrorString, "ncclGetErrorString");
#undef PQP_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

int need_nccl() {
    if (nccl().ok) return PQP_OK;
    pqp_set_err("no usable NCCL library (libnccl.so.2) could be loaded: %s", dlerror() ? "dlopen failed" : "symbols missing");
    return PQP_ERR_UNSUPPORTED;
}

#define PQP_NCCL(call)                                                             \
    do {                                                                           \
        ncclResult_t r_ = (call);                                                  \
        if (r_ != ncclSuccess) {                                                   \
            pqp_set_err("%s failed: %s", #call, nccl().GetErrorString(r_));        \
            return PQP_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

}  // namespace

struct pqp_multi {
    int n = 0;
    std::vector<pqp_handle *> h;
    std::vector<ncclComm_t> comm;
    std::vector<double *> d_gather;
    std::vector<size_t> gather_cap;
    std::vector<cudaEvent_t> ev_g0, ev_g1;
    std::vector<int> first_path, n_paths;
    std::vector<int64_t> first_station, n_stations;
    int64_t rows = 0;
};

extern "C" {

int pqp_nccl_unique_id(void *id) {
    if (!id) return PQP_ERR_ARG;
    int rc = need_nccl();
    if (rc != PQP_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == PQP_NCCL_ID_BYTES, "unique id size");
    ncclUniqueId u;
    PQP_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return PQP_OK;
}

int pqp_comm_init_rank(pqp_handle *h, int n_ranks, int rank, const void *id) {
    if (!h || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { pqp_set_err("pqp_comm_init_rank: bad argument"); return PQP_ERR_ARG; }
    int rc = need_nccl();
    if (rc != PQP_OK) return rc;
    if (h->nccl_comm) pqp_comm_destroy(h);
    PQP_CUDA(cudaSetDevice(h->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c;
    PQP_NCCL(nccl().CommInitRank(&c, n_ranks, u, rank));
    h->nccl_comm = (void *)c;
    return PQP_OK;
}

int pqp_comm_destroy(pqp_handle *h) {
    if (!h) return PQP_ERR_ARG;
    if (h->nccl_comm && nccl().ok) {
        cudaSetDevice(h->device);
        nccl().CommDestroy((ncclComm_t)h->nccl_comm);
    }
    h->nccl_comm = nullptr;
    return PQP_OK;
}

int pqp_allgather(pqp_handle *h, const double *d_send, double *d_recv, int64_t count, void *stream) {
    if (!h || !d_send || !d_recv || count < 0) { pqp_set_err("pqp_allgather: bad argument"); return PQP_ERR_ARG; }
    if (!h->nccl_comm) { pqp_set_err("pqp_allgather: no communicator (pqp_comm_init_rank first)"); return PQP_ERR_ARG; }
    PQP_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    PQP_NCCL(nccl().AllGather(d_send, d_recv, (size_t)count, ncclDouble, (ncclComm_t)h->nccl_comm, st));
    return PQP_OK;
}

void pqp_multi_destroy(pqp_multi *m) {
    if (!m) return;
    for (int d = 0; d < (int)m->h.size(); ++d) {
        if (!m->h[d]) continue;
        cudaSetDevice(m->h[d]->device);
        if (d < (int)m->comm.size() && m->comm[d] && nccl().ok) nccl().CommDestroy(m->comm[d]);
        if (d < (int)m->d_gather.size()) cudaFree(m->d_gather[d]);
        if (d < (int)m->ev_g0.size() && m->ev_g0[d]) cudaEventDestroy(m->ev_g0[d]);
        if (d < (int)m->ev_g1.size() && m->ev_g1[d]) cudaEventDestroy(m->ev_g1[d]);
        pqp_destroy(m->h[d]);
    }
    delete m;
}

int pqp_multi_create(pqp_multi **out, const pqp_params *params, int n_devices, const int *devices,
                     int max_batch_per_device, int max_total_points_per_device) {
    if (!out || !params || n_devices < 1 || !devices) { pqp_set_err("pqp_multi_create: bad argument"); return PQP_ERR_ARG; }
    *out = nullptr;
    int rc = need_nccl();
    if (rc != PQP_OK) return rc;
    pqp_multi *m = new (std::nothrow) pqp_multi;
    if (!m) return PQP_ERR_ARG;
    m->n = n_devices;
    m->h.assign(n_devices, nullptr);
    m->comm.assign(n_devices, nullptr);
    m->d_gather.assign(n_devices, nullptr);
    m->gather_cap.assign(n_devices, 0);
    m->ev_g0.assign(n_devices, nullptr);
    m->ev_g1.assign(n_devices, nullptr);
    m->first_path.assign(n_devices, 0); m->n_paths.assign(n_devices, 0);
    m->first_station.assign(n_devices, 0); m->n_stations.assign(n_devices, 0);
    for (int d = 0; d < n_devices; ++d) {
        rc = pqp_create(&m->h[d], params, devices[d], max_batch_per_device, max_total_points_per_device);
        if (rc != PQP_OK) { pqp_multi_destroy(m); return rc; }
        cudaSetDevice(devices[d]);
        if (cudaEventCreate(&m->ev_g0[d]) != cudaSuccess || cudaEventCreate(&m->ev_g1[d]) != cudaSuccess) {
            pqp_set_err("cudaEventCreate failed");
            pqp_multi_destroy(m);
            return PQP_ERR_CUDA;
        }
    }
    ncclResult_t r = nccl().CommInitAll(m->comm.data(), n_devices, devices);
    if (r != ncclSuccess) {
        pqp_set_err("ncclCommInitAll failed: %s", nccl().GetErrorString(r));
        for (auto &c : m->comm) c = nullptr;
        pqp_multi_destroy(m);
        return PQP_ERR_CUDA;
    }
    *out = m;
    return PQP_OK;
}

int pqp_multi_devices(const pqp_multi *m) { return m ? m->n : 0; }
const double *pqp_multi_gathered(const pqp_multi *m, int k) { return (m && k >= 0 && k < m->n) ? m->d_gather[k] : nullptr; }
int64_t pqp_multi_gather_rows(const pqp_multi *m) { return m ? m->rows : 0; }
int pqp_multi_shard(const pqp_multi *m, int k, int *first_path, int *n_paths, int64_t *first_station) {
    if (!m || k < 0 || k >= m->n) return PQP_ERR_ARG;
    if (first_path) *first_path = m->first_path[k];
    if (n_paths) *n_paths = m->n_paths[k];
    if (first_station) *first_station = m->first_station[k];
    return PQP_OK;
}

int pqp_multi_solve_batch(pqp_multi *m, int formulation, int batch, const int32_t *n_points, const pqp_state *ref,
                          const pqp_station_bounds *bounds, const double *x0, const double *end_heading,
                          pqp_state *out_states, double *out_frenet, int32_t *status, int32_t *iters, int gather,
                          pqp_stats *stats) {
    if (!m || batch < 0 || (batch > 0 && (!n_points || !ref || !bounds || !x0 || !end_heading || !out_states || !status))) {
        pqp_set_err("pqp_multi_solve_batch: bad argument");
        return PQP_ERR_ARG;
    }
    if (formulation != PQP_FORM_KP && formulation != PQP_FORM_K) {   // (KPC would need the per-station limits as well)
        pqp_set_err("pqp_multi_solve_batch: KP and K only");
        return PQP_ERR_UNSUPPORTED;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (batch == 0) return PQP_OK;
    // contiguous shards of (nearly) equal station count
    std::vector<int64_t> off((size_t)batch + 1, 0);
    for (int b = 0; b < batch; ++b) {
        if (n_points[b] < 0) { pqp_set_err("negative n_points"); return PQP_ERR_ARG; }
        off[b + 1] = off[b] + n_points[b];
    }
    const int64_t total = off[batch];
    int b0 = 0;
    m->rows = 0;
    for (int d = 0; d < m->n; ++d) {
        const int64_t target = total * (d + 1) / m->n;
        int b1 = b0;
        while (b1 < batch && (off[b1 + 1] <= target || d == m->n - 1)) ++b1;
        if (d == m->n - 1) b1 = batch;
        m->first_path[d] = b0; m->n_paths[d] = b1 - b0;
        m->first_station[d] = off[b0]; m->n_stations[d] = off[b1] - off[b0];
        m->rows = std::max(m->rows, m->n_stations[d]);
        if (b1 - b0 > m->h[d]->max_batch || off[b1] - off[b0] > m->h[d]->max_total) {
            pqp_set_err("a shard exceeds the per-device capacity the solver was created with");
            return PQP_ERR_CAPACITY;
        }
        b0 = b1;
    }
    std::vector<int32_t> iters_local;
    if (!iters) { iters_local.resize((size_t)batch); iters = iters_local.data(); }
    // (1) enqueue upload + kernels + download of every shard; nothing is synchronised yet
    int rc = PQP_OK;
    for (int d = 0; d < m->n && rc == PQP_OK; ++d) {
        if (!m->n_paths[d]) continue;
        pqp_handle *h = m->h[d];
        const int pb = m->first_path[d];
        const int64_t sb = m->first_station[d];
        h->defer_sync = true;
        h->force_frenet = gather != 0;
        rc = pqp_solve_batch(h, formulation, m->n_paths[d], n_points + pb, ref + sb, bounds + sb, x0 + 3 * (size_t)pb,
                             end_heading + pb, nullptr, nullptr, out_states + sb, out_frenet ? out_frenet + 3 * sb : nullptr,
                             status + pb, iters + pb, nullptr);
        h->defer_sync = false;
        h->force_frenet = false;
    }
    // (2) one all-gather of the Frenet states: every device contributes `rows` stations (its shard, padded)
    if (rc == PQP_OK && gather) {
        const size_t count = (size_t)m->rows * 3;
        for (int d = 0; d < m->n && rc == PQP_OK; ++d) {
            cudaSetDevice(m->h[d]->device);
            if (m->gather_cap[d] < count * m->n) {
                cudaStreamSynchronize(m->h[d]->stream);
                cudaFree(m->d_gather[d]);
                m->d_gather[d] = nullptr; m->gather_cap[d] = 0;
                if (cudaMalloc(&m->d_gather[d], count * m->n * sizeof(double)) != cudaSuccess) {
                    pqp_set_err("cudaMalloc of the gather buffer failed");
                    rc = PQP_ERR_CUDA;
                    break;
                }
                m->gather_cap[d] = count * m->n;
            }
            if ((int64_t)m->h[d]->max_total < m->rows) { pqp_set_err("gather rows exceed the per-device capacity"); rc = PQP_ERR_CAPACITY; }
        }
        if (rc == PQP_OK) {
            ncclResult_t r = nccl().GroupStart();
            for (int d = 0; d < m->n && r == ncclSuccess; ++d) {
                cudaSetDevice(m->h[d]->device);
                cudaEventRecord(m->ev_g0[d], m->h[d]->stream);
                // (a device whose shard is empty or short contributes whatever its padded rows hold: zeros from cudaMalloc or old results)
                r = nccl().AllGather(m->h[d]->d_frenet, m->d_gather[d], count, ncclDouble, m->comm[d], m->h[d]->stream);
            }
            ncclResult_t r2 = nccl().GroupEnd();
            if (r == ncclSuccess) r = r2;
            for (int d = 0; d < m->n; ++d) {
                cudaSetDevice(m->h[d]->device);
                cudaEventRecord(m->ev_g1[d], m->h[d]->stream);
            }
            if (r != ncclSuccess) { pqp_set_err("ncclAllGather failed: %s", nccl().GetErrorString(r)); rc = PQP_ERR_CUDA; }
        }
    }
    // (3) drain every device (also after an error: copies into the caller's buffers may be in flight)
    for (int d = 0; d < m->n; ++d) {
        cudaSetDevice(m->h[d]->device);
        if (cudaStreamSynchronize(m->h[d]->stream) != cudaSuccess && rc == PQP_OK) { pqp_set_err("cudaStreamSynchronize failed"); rc = PQP_ERR_CUDA; }
        cudaStreamSynchronize(m->h[d]->stream2);
    }
    if (rc != PQP_OK) return rc;
    if (stats) {
        for (int d = 0; d < m->n; ++d) {
            if (!m->n_paths[d]) continue;
            pqp_handle *h = m->h[d];
            cudaSetDevice(h->device);
            float up = 0, kern = 0, down = 0, g = 0;
            cudaEventElapsedTime(&up, h->ev[0], h->ev[1]);
            cudaEventElapsedTime(&kern, h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&down, h->ev[2], h->ev[3]);
            if (gather) cudaEventElapsedTime(&g, m->ev_g0[d], m->ev_g1[d]);
            stats->h2d_ms = std::max(stats->h2d_ms, up);
            stats->kernel_ms = std::max(stats->kernel_ms, kern);
            stats->d2h_ms = std::max(stats->d2h_ms, down + g);
            stats->kernel_launches += h->deferred_launches;
        }
        for (int b = 0; b < batch; ++b) {
            stats->total_iters += iters[b];
            stats->max_iters = std::max(stats->max_iters, iters[b]);
            stats->n_solved += (status[b] == PQP_SOLVED);
        }
        stats->h2d_bytes = (int64_t)(total * (sizeof(pqp_state) + sizeof(pqp_station_bounds)) + (int64_t)batch * 44);
        stats->d2h_bytes = (int64_t)(total * (sizeof(pqp_state) + (out_frenet ? 24 : 0)) + (int64_t)batch * 8);
    }
    return PQP_OK;
}

}  // extern "C"
