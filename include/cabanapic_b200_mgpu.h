/*
 * cabanapic_b200 -- multi-GPU layer of the C ABI: one process per GPU, NCCL over NVLink / NVSwitch.
 *
 * The reference is single-process: its only trace of a decomposition is the `contribute` call where a
 * distributed run would reduce the accumulator (example/example.cpp:248-251, with the
 * "// TODO: boundaries? MPI" line at :254) and the commented VPIC neighbour logic of
 * src/move_p.h:327-346.  SURVEY.md 8(b)/(e) lists what a replacement must export on top of
 * cabanapic_b200.h: create(..., ngpus, mode), reduce_accumulator, migration_counts.  This is that.
 *
 * Two modes:
 *   REPLICATED  every rank holds the whole grid and 1/world of the particles; the accumulators are
 *               summed with ncclAllReduce exactly where the reference calls `contribute`; every rank
 *               then runs the identical field solve.
 *   SLAB        z-slabs (a z-plane incl. its x/y ghosts is contiguous under VOXEL, src/types.h:195);
 *               per step the ranks exchange with their -z / +z neighbours (periodic ring) the ghost
 *               accumulator planes + the particles that crossed a z face (one NCCL group, every count
 *               on the device), the z sweeps of the J fold (src/fields.h:126-183), the J ghost-copy
 *               planes (:33-98) and the cB ghost-copy planes after each advance_b (:718).  Copy planes
 *               are received straight into the field arrays; no host synchronisation inside a step;
 *               pairs of steps can be replayed from a CUDA graph (NCCL kernels included).
 * NCCL is loaded at run time (dlopen "libnccl.so.2": the copy a host framework such as torch already
 * loaded, else the system one), so the library has no link-time dependency on it.
 *
 * Rendezvous is the caller's business: rank 0 obtains 128 opaque bytes from cpic_mgpu_unique_id and
 * hands them to the other ranks (MPI_Bcast, a torch.distributed store, or -- for launchers that have
 * nothing -- cpic_mgpu_bootstrap_file, which uses a file all ranks can see).
 */
#ifndef CABANAPIC_B200_MGPU_H
#define CABANAPIC_B200_MGPU_H

#include "cabanapic_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { CPIC_MGPU_REPLICATED = 0, CPIC_MGPU_SLAB = 1, CPIC_MGPU_AUTO = 2 };
#define CPIC_MGPU_ID_BYTES 128

typedef struct cpic_mgpu cpic_mgpu;

const char* cpic_mgpu_last_error(const cpic_mgpu* m);   /* m may be NULL: error of the last failed create / bootstrap */

/* ncclGetUniqueId on the calling rank (rank 0). */
int  cpic_mgpu_unique_id(void* id_out /* CPIC_MGPU_ID_BYTES */);
/* File rendezvous for launchers without a broadcast of their own: rank 0 creates the id and writes it to
 * `path` (atomically, via rename), the others wait for the file (up to timeout_s seconds) and read it. */
int  cpic_mgpu_bootstrap_file(const char* path, int32_t rank, int32_t world, double timeout_s, void* id_out);

/* `global` describes the WHOLE box (nx, ny, nz = global interior cells; max_particles = capacity of THIS rank's
 * store; device = this rank's CUDA device).  mode AUTO picks SLAB when every rank gets >= 2 planes of a grid of
 * >= 2^18 cells, else REPLICATED.  SLAB needs float, the EM solver and enable_sort (the reordering push);
 * send_capacity = particles per direction and step the migration buffers hold (0: 5 % of a plane's particles at
 * max_particles' density, at least 4096).  Collective: every rank of the communicator must call it. */
int  cpic_mgpu_create(const cpic_params* global, int32_t rank, int32_t world, const void* unique_id, int32_t mode,
                      int64_t send_capacity, cpic_mgpu** out);
void cpic_mgpu_destroy(cpic_mgpu* m);

/* This rank's context (every cabanapic_b200.h entry point works on it: uploads, downloads, diagnostics) and layout:
 * z0 = first global interior plane owned here, nzl = number of planes (REPLICATED: 0, nz). */
cpic_ctx* cpic_mgpu_context(cpic_mgpu* m);
int  cpic_mgpu_layout(const cpic_mgpu* m, int32_t* mode, int32_t* z0, int32_t* nzl);

/* How the slab exchanges travel.  PEER_MEMORY (the default when the ranks can map each other's memory through CUDA
 * IPC): one kernel stores the boundary planes / leaver records straight into the z neighbours' memory over NVLink /
 * NVSwitch and raises their arrival flags; NCCL: ncclSend / ncclRecv groups (fallback; CPIC_MGPU_P2P=0 forces it);
 * NONE: a single rank.  CPIC_P2P_TIMEOUT_S (default 120) bounds the wait for a neighbour. */
enum { CPIC_MGPU_TRANSPORT_NONE = 0, CPIC_MGPU_TRANSPORT_NCCL = 1, CPIC_MGPU_TRANSPORT_PEER_MEMORY = 2 };
int  cpic_mgpu_transport(const cpic_mgpu* m);

/* This rank's share of the synthetic uniform plasma of cpic_init_uniform_plasma over the GLOBAL box:
 * SLAB: the particles of the planes it owns; REPLICATED: global particles [N*rank/world, N*(rank+1)/world). */
int  cpic_mgpu_init_uniform_plasma(cpic_mgpu* m, int32_t nppc, uint64_t seed, double vthx, double vthy, double vthz,
                                   double weight);

/* Kokkos::Experimental::contribute of a distributed run (example/example.cpp:248-251): REPLICATED: ncclAllReduce(sum)
 * of the accumulator; SLAB: the ghost-plane exchange of the accumulator (without the particle migration). */
int  cpic_mgpu_reduce_accumulator(cpic_mgpu* m);

/* nsteps whole steps of example/example.cpp:221-266 with the exchanges woven in.  sort_interval as cpic_step
 * (SLAB supports CPIC_SORT_FUSED only).  use_graph != 0 (SLAB): capture a pair of steps once and replay it
 * (falls back to eager launches if the capture fails).  No host synchronisation. */
int  cpic_mgpu_step(cpic_mgpu* m, const cpic_consts* k, int64_t nsteps, int32_t sort_interval, int32_t use_graph);

/* Capture the CUDA graph of a pair of steps now, without executing anything (SLAB, after at least two eager steps have
 * brought every lazily allocated buffer and the device-side particle count into being); cpic_mgpu_step(use_graph)
 * otherwise captures on its first eligible call.  Returns CPIC_E_UNSUPPORTED when the graph path does not apply. */
int  cpic_mgpu_prepare_graph(cpic_mgpu* m, const cpic_consts* k);

/* One step for a caller whose slab lives in HOST memory (every array of the reference's host build does,
 * example/example.cpp:58-113) -- cpic_step_host for the slab mode.  in[8] / out[8]: this rank's particle members
 * (dx dy dz ux uy uz w cell, cell indices in this slab's numbering, interior planes only), n of them; fields_in[9] /
 * fields_out[9]: this slab's field members incl. ghosts.  The particles stream through the device in chunks (H2D of
 * chunk i+1, in-place push of chunk i, D2H of chunk i-1 overlap), the ghost-plane / leaver exchanges and the field
 * advance follow, and `out` is then patched where the migration changed the store: holes left by the leavers are
 * filled from the tail, arrivals appended.  *n_out = this rank's particle count after the step (<= out_capacity).
 * Particle order is the caller's otherwise.  SLAB mode only. */
int  cpic_mgpu_step_host(cpic_mgpu* m, const cpic_consts* k, const void* const in[8], void* const out[8], int64_t n,
                         int64_t out_capacity, int64_t* n_out, const void* const fields_in[9], void* const fields_out[9]);

/* Particles this rank sent to its lower / upper neighbour since creation (SLAB; synchronises). */
int  cpic_mgpu_migration_counts(cpic_mgpu* m, int64_t out[2]);
/* ... and in the last step only. */
int  cpic_mgpu_last_migration(cpic_mgpu* m, int64_t out[2]);

/* Global diagnostics (collective; every rank receives the result):
 * energies: cpic_energies summed over the slabs (REPLICATED: every rank already holds the total);
 * digest: out[0] particles, [1] sum of weights, [2] particles whose cell is not an interior voxel of their rank,
 *         [3] particles with an offset outside [-1,1], [4] kinetic energy (cpic_kinetic_energy), [5] field energy E,
 *         [6] field energy B, [7] particles migrated (sent down + up) so far -- all summed over the ranks. */
int  cpic_mgpu_energies(cpic_mgpu* m, double* e_energy, double* b_energy);
int  cpic_mgpu_state_digest(cpic_mgpu* m, double out[8]);

int  cpic_mgpu_sync(cpic_mgpu* m);
/* 1 if the last cpic_mgpu_step replayed a CUDA graph */
int  cpic_mgpu_used_graph(const cpic_mgpu* m);

#ifdef __cplusplus
}
#endif
#endif /* CABANAPIC_B200_MGPU_H */
