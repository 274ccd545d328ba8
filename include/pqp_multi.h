/*
 * pqp_multi.h -- multi-GPU entry points of libpqp.so: the batch shards over the GPUs of one box with no data-path
 * collective (paths are independent) and ONE NCCL all-gather of the solved Frenet states at the end
 * (BASELINE.json north_star).  Host orchestration is C++ inside the library; NCCL is resolved at run time
 * (dlopen of libnccl.so.2 -- the copy already loaded into the process, e.g. PyTorch's, else the system one), so
 * libpqp.so has no link-time dependency on it and single-GPU users never touch it.
 *
 * Two ways to use it:
 *   (a) one process per GPU (torchrun / MPI): pqp_nccl_unique_id on rank 0, broadcast the 128 bytes by any transport,
 *       pqp_comm_init_rank on every rank's handle, then pqp_allgather after each solve;
 *   (b) one process driving several GPUs: pqp_multi_create / pqp_multi_solve_batch.
 * No reference counterpart: the reference plans one path at a time on one CPU thread.
 */
#ifndef PQP_MULTI_H_
#define PQP_MULTI_H_
#include "pqp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PQP_NCCL_ID_BYTES 128

/* (a) one process per GPU ------------------------------------------------------------------------------------ */

/* Fill id[128] with a fresh NCCL unique id (rank 0; the caller distributes it). */
int pqp_nccl_unique_id(void *id);

/* Join the communicator of `n_ranks` processes as `rank`, bound to the handle's device.  Collective: every rank calls
 * it with the same id.  PQP_ERR_UNSUPPORTED when no NCCL library can be loaded. */
int pqp_comm_init_rank(pqp_handle *h, int n_ranks, int rank, const void *id);

/* All-gather `count` doubles per rank from d_send into d_recv ([n_ranks * count], rank order) on `stream`
 * (NULL: the handle's stream); asynchronous.  Used for the solved Frenet states: count = 3 * (padded) stations. */
int pqp_allgather(pqp_handle *h, const double *d_send, double *d_recv, int64_t count, void *stream);

/* Leave the communicator (also done by pqp_destroy). */
int pqp_comm_destroy(pqp_handle *h);

/* (b) one process, several GPUs ------------------------------------------------------------------------------- */

typedef struct pqp_multi pqp_multi;

/* One solver per listed device + one communicator over them (ncclCommInitAll).  Capacities are per device. */
int pqp_multi_create(pqp_multi **out, const pqp_params *params, int n_devices, const int *devices,
                     int max_batch_per_device, int max_total_points_per_device);
void pqp_multi_destroy(pqp_multi *m);
int pqp_multi_devices(const pqp_multi *m);

/* pqp_solve_batch over all devices: the batch is cut into contiguous shards of (nearly) equal station count, every
 * device uploads, solves and downloads its shard concurrently (host buffers as in pqp_solve_batch; "KP" and "K"), and, when
 * `gather` is non-zero, ONE all-gather leaves the Frenet states of the WHOLE batch on every device:
 * pqp_multi_gathered(m, d) = device pointer on device index d to [n_devices][rows][3] doubles, rows =
 * pqp_multi_gather_rows(m) (the largest shard's station count; shard k's stations start at row 0 of block k).
 * pqp_multi_shard reports shard k = paths [first_path, first_path + n_paths) / stations from first_station on.
 * stats (optional): kernel_ms = slowest device's kernel span, d2h_ms = the all-gather on the slowest device. */
int pqp_multi_solve_batch(pqp_multi *m, int formulation, int batch, const int32_t *n_points, const pqp_state *ref,
                          const pqp_station_bounds *bounds, const double *x0, const double *end_heading,
                          pqp_state *out_states, double *out_frenet, int32_t *status, int32_t *iters,
                          int gather, pqp_stats *stats);
const double *pqp_multi_gathered(const pqp_multi *m, int device_index);
int64_t pqp_multi_gather_rows(const pqp_multi *m);
int pqp_multi_shard(const pqp_multi *m, int k, int *first_path, int *n_paths, int64_t *first_station);

#ifdef __cplusplus
}
#endif
#endif /* PQP_MULTI_H_ */
