/*
 * cabanapic_b200 -- C ABI of the B200-native CabanaPIC particle hot path.
 *
 * The reference (ECP-copa/CabanaPIC) has no FFI layer: its hot path sits behind
 * plain C++ free functions called from example/example.cpp.  Every entry point
 * below replaces one of those call sites (cited as path:line relative to the
 * reference tree) and is what a reference-side binding would call; see
 * INTEGRATION.md for the stub a maintainer would add.
 *
 * Conventions
 *   - opaque context, one host thread per context, one CUDA device per context;
 *   - every function returns 0 on success or a negative CPIC_E_* code;
 *     cpic_last_error(ctx) gives the message (no exceptions cross the boundary);
 *   - work is enqueued in order on the context's stream; only download_*,
 *     energies, migration/query calls and cpic_sync block the host;
 *   - `real` buffers are float or double according to cpic_params.real_bytes
 *     (the reference's -DREAL_TYPE, src/types.h:4-8), passed as void*;
 *   - particle members are the reference's AoSoA members (src/types.h:31-58):
 *     dx dy dz (cell-local offset in [-1,1]), ux uy uz (momentum), w, cell
 *     (int voxel index incl. ghosts, VOXEL() of src/types.h:195);
 *   - field members in FieldFields order (src/types.h:138-149):
 *     ex ey ez cbx cby cbz jfx jfy jfz, each num_cells long;
 *   - interpolators are exchanged as [cell][18] (src/types.h:64-84),
 *     accumulators as [cell][3][4] (src/types.h:116-120).
 * There is no CPU fallback: creating a context without a usable CUDA device
 * fails with CPIC_E_CUDA.
 */
#ifndef CABANAPIC_B200_H
#define CABANAPIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPIC_ABI_VERSION 1

enum {
    CPIC_OK = 0,
    CPIC_E_INVALID = -1,     /* bad argument / unsupported configuration     */
    CPIC_E_CUDA = -2,        /* CUDA runtime error (message has the detail)  */
    CPIC_E_NOMEM = -3,       /* device allocation failed                     */
    CPIC_E_CAPACITY = -4,    /* more particles than max_particles            */
    CPIC_E_BAD_CELL = -5,    /* particle cell index outside the grid         */
    CPIC_E_UNSUPPORTED = -6  /* a combination that is not implemented        */
};

enum { CPIC_SOLVER_EM = 0, CPIC_SOLVER_ES_1D = 1 };   /* -DSOLVER_TYPE, example/example.cpp:24-35 */
/* enum Boundary, src/input/deck.h:9-12.  PERIODIC is what the reference runs.  REFLECT is declared there but exits in the
 * field code (src/fields.h:21-25,113-117) and survives in the mover only as VPIC's commented block (src/move_p.h:298-324);
 * here it is built: a particle whose streak ends on a domain face keeps its cell, stays on the face and has
 * its momentum component and remaining displacement along that axis reversed; the field side is a perfectly conducting box
 * (E_tang = 0 on the six walls, "anti_symmetric_fields" of src/grid.h:4-17; no ghost fold / copy).  EM solver, single
 * context or replicated multi-GPU mode.  No reference output exists for it: parity is pinned by oracle/cpic_oracle.c only. */
enum { CPIC_BOUNDARY_REFLECT = 0, CPIC_BOUNDARY_PERIODIC = 1 };

/* Floating-point policy of the particle kernels.
 * STRICT: every multiply/add rounds separately, IEEE sqrt/div -- bit-identical
 *         per-particle results to the reference's CPU build (gcc -O2, no FMA).
 * CONTRACT: a*b+c chains are fused (what nvcc does to the reference's Kokkos-CUDA
 *         build by default); results agree to rounding, not bitwise. */
enum { CPIC_FP_STRICT = 0, CPIC_FP_CONTRACT = 1 };

/* How push() adds the 12 per-streak currents into the accumulator.
 * All modes give the same sums up to floating-point association. */
enum {
    CPIC_DEPOSIT_AUTO = 0,        /* = WARP                                               */
    CPIC_DEPOSIT_ATOMIC = 1,      /* 12 scalar global atomics per streak                  */
    CPIC_DEPOSIT_ATOMIC_V4 = 2,   /* 3 x 128-bit vector atomics per streak (float only)   */
    CPIC_DEPOSIT_WARP = 3,        /* warp-aggregated per cell, then one atomic row/warp   */
    CPIC_DEPOSIT_ORDERED = 4      /* parity runs: every streak is recorded and the streaks of a cell are added in (particle,
                                     streak) order -- the summation order of the reference's serial loop (src/push.h:218-254,
                                     src/move_p.h:154-190), so accumulators, J and whole histories are bit-identical to it in
                                     strict FP mode.  In-place push only (the particle order must be the reference's), at most
                                     2^22 particles, not fast.                            */
};

typedef struct cpic_ctx cpic_ctx;

typedef struct cpic_params {
    int32_t nx, ny, nz, ng;       /* interior cells per axis; ng must be 1 (src/move_p.h:19-47 hard-wires it) */
    int32_t real_bytes;           /* 4 or 8 */
    int32_t solver;               /* CPIC_SOLVER_* */
    int32_t boundary;             /* CPIC_BOUNDARY_* */
    int32_t device;               /* CUDA device ordinal */
    int32_t fp_mode;              /* CPIC_FP_* */
    int32_t deposit_mode;         /* CPIC_DEPOSIT_* */
    int64_t max_particles;        /* capacity of the particle store */
    int32_t enable_sort;          /* allocate the second particle buffer needed by cpic_sort_particles */
    int32_t reserved[7];
} cpic_params;

/* Step constants exactly as the reference driver derives them in real_t
 * (example/example.cpp:77-113, 179-181), widened to double for transport. */
typedef struct cpic_consts {
    double qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
    double dx, dy, dz, dt;
    double px, py, pz, dt_eps0;
} cpic_consts;

/* Diagnostics of the last push (optional). */
typedef struct cpic_push_stats {
    int64_t movers;     /* particles that left their cell (took the move_p path) */
    int64_t crossings;  /* total cell faces crossed                               */
    int64_t wraps[6];   /* periodic wraps per face: -x -y -z +x +y +z              */
} cpic_push_stats;

int  cpic_abi_version(void);
const char* cpic_last_error(const cpic_ctx* ctx);   /* ctx may be NULL: error of the last failed cpic_create */

/* particle_list_t / field_array_t / interpolator_array_t / accumulator_array_t
 * allocation, example/example.cpp:121-144; Field_Solver ctor zeroing, src/fields.h:279-315;
 * initialize_interpolator, src/interpolator.cpp:125-172. */
int  cpic_create(const cpic_params* params, cpic_ctx** out);
void cpic_destroy(cpic_ctx* ctx);
/* Multiple species: the reference is single-species through the scalars qsp / me of the deck
 * (src/input/deck.h:254-261); the VPIC decks it ships show the intended form, one particle list per species pushed
 * with its own charge and mass into the same accumulator (decks/vpic/2stream-em0.cxx:207-208).  cpic_create_species
 * gives `parent` (a context of cpic_create) another particle store with its own capacity, sort state and push
 * constants; it SHARES the parent's field, interpolator and accumulator arrays and its stream, so every entry point of
 * this header works on it -- uploads / downloads of ITS particles, cpic_push / cpic_push_reorder with ITS cpic_consts
 * (qsp, qdt_2mc = qsp dt / (2 m c)) deposit into the common J -- while the field-side calls belong to the parent.
 * Destroy the species before the parent.  cpic_step_species: nsteps of example/example.cpp:221-266 with the push
 * repeated for every listed species (the parent itself may be one of them) between clear and unload; consts[s] belongs
 * to species[s], the field constants are taken from consts[0]. */
int  cpic_create_species(cpic_ctx* parent, int64_t max_particles, cpic_ctx** out);
int  cpic_step_species(cpic_ctx* ctx, cpic_ctx* const* species, const cpic_consts* consts, int32_t nspecies, int64_t nsteps,
                       int32_t sort_interval, double* energies);
int  cpic_sync(cpic_ctx* ctx);
int  cpic_num_cells(const cpic_ctx* ctx, int64_t* out);
int  cpic_num_particles(const cpic_ctx* ctx, int64_t* out);

/* Deck initialisers run on the host (src/input/deck.h:209-252); these move the result. */
int  cpic_upload_particles(cpic_ctx* ctx, const void* dx, const void* dy, const void* dz, const void* ux,
                           const void* uy, const void* uz, const void* w, const int32_t* cell, int64_t n);
int  cpic_download_particles(cpic_ctx* ctx, void* dx, void* dy, void* dz, void* ux, void* uy, void* uz, void* w,
                             int32_t* cell, int64_t capacity, int64_t* n_out);
int  cpic_upload_fields(cpic_ctx* ctx, const void* const fields[9]);
int  cpic_download_fields(cpic_ctx* ctx, void* const fields[9]);
int  cpic_upload_interpolators(cpic_ctx* ctx, const void* interp /* [nc][18] */);
int  cpic_download_interpolators(cpic_ctx* ctx, void* interp /* [nc][18] */);
int  cpic_upload_accumulators(cpic_ctx* ctx, const void* acc /* [nc][12] */);
int  cpic_download_accumulators(cpic_ctx* ctx, void* acc /* [nc][12] */);

/* The time-loop call surface, example/example.cpp:221-266 */
int  cpic_load_interpolator_array(cpic_ctx* ctx);                 /* src/interpolator.cpp:4-124   */
int  cpic_initialize_interpolator(cpic_ctx* ctx);                 /* src/interpolator.cpp:125-172 */
int  cpic_clear_accumulator_array(cpic_ctx* ctx);                 /* src/accumulator.cpp:5-43     */
int  cpic_push(cpic_ctx* ctx, const cpic_consts* k);              /* src/push.h:7-300 + src/move_p.h:59-374 */
int  cpic_contribute(cpic_ctx* ctx);                              /* Kokkos contribute/reset_except, example.cpp:248-251 (no-op on one GPU) */
int  cpic_unload_accumulator_array(cpic_ctx* ctx, const cpic_consts* k); /* src/accumulator.cpp:45-122 */
int  cpic_advance_b(cpic_ctx* ctx, double px, double py, double pz);     /* src/fields.h:352-364, 668-719 (ES: no-op) */
int  cpic_advance_e(cpic_ctx* ctx, double px, double py, double pz, double dt_eps0); /* src/fields.h:365-378, 618-665, 511-544 */
int  cpic_uncenter_particles(cpic_ctx* ctx, double qdt_2mc);      /* src/uncenter_p.h:4-105       */
int  cpic_energies(cpic_ctx* ctx, double* e_energy, double* b_energy);   /* src/fields.h:556-615, 484-509 */
/* Kinetic energy of the particles, sum_p w_p (gamma_p - 1) with gamma = sqrt(1 + u.u), accumulated in double
 * (in units of m c^2 per unit weight).  The reference has no such diagnostic; together with
 * cpic_energies it closes the energy budget: w_p (gamma-1) m c^2 + (eps0/2) sum (E^2 + cB^2) dV is conserved by
 * the scheme up to grid heating, a size-independent check of the whole loop. */
int  cpic_kinetic_energy(cpic_ctx* ctx, double* out);
/* One-pass digest of the particle store + the field energies (bench / multi-GPU parity block): out[0] particles,
 * [1] sum of weights, [2] particles whose cell is not an interior voxel, [3] particles with an offset outside [-1,1],
 * [4] kinetic energy (as cpic_kinetic_energy), [5] E energy, [6] B energy (as cpic_energies), [7] 0 (reserved: particles
 * migrated, see cpic_mgpu_state_digest). */
int  cpic_state_digest(cpic_ctx* ctx, double out[8]);
int  cpic_update_ghosts(cpic_ctx* ctx, int which /* 0: fold J, 1: copy J, 2: copy cB, 3 / 4: first / second sweep of the J fold only */); /* src/fields.h:11-271 */

/* n whole steps in the reference's order (example/example.cpp:221-266), fused on the
 * device: no host synchronisation between steps.  sort_interval > 0 re-sorts the particles
 * by cell every that many steps (the step the reference left commented out,
 * example/example.cpp:224-228).  energies (optional, host) receives (e,b) after every step. */
#define CPIC_SORT_FUSED (-1)   /* sort_interval: keep the store cell-ordered with cpic_push_reorder, no sort pass */
int  cpic_step(cpic_ctx* ctx, const cpic_consts* k, int64_t nsteps, int32_t sort_interval, double* energies);

/* ONE step of the same loop for a caller whose state lives in HOST memory (the reference's host build
 * keeps everything there: example/example.cpp:58-113 allocates, :221-266 steps).  Equivalent to
 * cpic_upload_particles + cpic_upload_fields + cpic_step(k, 1, 0, energies) + cpic_download_particles +
 * cpic_download_fields, but the particles STREAM through the device in chunks: chunk i is pushed (in place,
 * caller's particle order kept) while chunk i+1 is still arriving and chunk i-1 is already on its way
 * back, so PCIe runs in both directions at once and the step costs max(H2D, D2H) instead of their sum.
 * The eight `in` member arrays (dx dy dz ux uy uz w cell) hold n particles; the eight `out` arrays receive
 * them after the step (out[m] may equal in[m]; out == NULL skips the download, an out[m] == NULL skips
 * that member).  fields_in / fields_out: nine arrays of cpic_num_cells reals, as cpic_upload_fields /
 * cpic_download_fields (fields_out may be NULL).  energies: NULL or 2 doubles (e, b after the step).
 * Pinned host memory is what makes the copies asynchronous; pageable memory works but serialises.
 * Afterwards the context holds the advanced state, as after the unfused sequence. */
int  cpic_step_host(cpic_ctx* ctx, const cpic_consts* k, const void* const in[8], void* const out[8], int64_t n,
                    const void* const fields_in[9], void* const fields_out[9], double* energies);

/* Device-side Particle_Initializer for the synthetic uniform thermal plasma (the reference's
 * default initialisers also run in the execution space, src/input/deck.h:155-157): fills this
 * context's store with global particles [first, first+count) of a gnx*gny*gnz*nppc box;
 * particle k sits in global interior cell k/nppc (x fastest), offsets U(-1,1), momenta
 * N(0,vth) from Philox-4x32-10(seed, k).  z0 = first global z-plane owned here (slab mode). */
int  cpic_init_uniform_plasma(cpic_ctx* ctx, int64_t first, int64_t count, int32_t gnx, int32_t gny, int32_t gnz,
                              int32_t nppc, int32_t z0, uint64_t seed, double vthx, double vthy, double vthz,
                              double weight);

/* Cabana::sortByKey by Cell_Index (example/example.cpp:224-228). */
int  cpic_sort_particles(cpic_ctx* ctx);
/* push<> (src/push.h:7-300) and that sortByKey in ONE pass over the particles: the advanced particles are
 * written to the second particle buffer, each into the segment of the cell it occupied when the call
 * began, so the store the next step reads is cell-ordered up to the particles that changed cell in this
 * one step; the call also produces the cell histogram the next call's segments come from.  Same
 * per-particle results as cpic_push; the order of the particles in the store changes (as with any sort).
 * float + CPIC_DEPOSIT_WARP/AUTO contexts with enable_sort; otherwise equal to cpic_sort_particles +
 * cpic_push. */
int  cpic_push_reorder(cpic_ctx* ctx, const cpic_consts* k);
int  cpic_enable_push_stats(cpic_ctx* ctx, int32_t on);   /* count movers/crossings/wraps in cpic_push (off by default) */
int  cpic_push_stats_get(cpic_ctx* ctx, cpic_push_stats* out);

/* Interop for host frameworks that own streams / collectives (torch.distributed, NCCL):
 * raw device pointers of the context's arrays.  which: 0..7 particle members,
 * 16 fields (9*nc_pad contiguous, member m at offset m*stride), 17 interpolators, 18 accumulators. */
int  cpic_device_ptr(cpic_ctx* ctx, int which, void** ptr, int64_t* count, int64_t* stride);
int  cpic_set_stream(cpic_ctx* ctx, void* cuda_stream);
int  cpic_set_num_particles(cpic_ctx* ctx, int64_t n);   /* after an external particle exchange */
int  cpic_set_modes(cpic_ctx* ctx, int32_t fp_mode, int32_t deposit_mode);

/* Multi-GPU z-slab decomposition (new: the reference is single-process).
 * cpic_set_axis_periodic clears the periodic wrap along chosen axes inside this context: the mover
 * then leaves a particle that crossed such a face in the ghost cell (the reference's behaviour for a
 * non-periodic boundary, src/move_p.h:257-352) and the ghost fold/copy kernels skip that axis (the
 * host exchanges those planes with the neighbouring slab).
 * cpic_extract_z_leavers removes the particles sitting in the ghost planes z = 0 (-> lo_buf) and
 * z = nz+1 (-> hi_buf) from the store, keeps the store dense, and adds rebase_lo / rebase_hi to their
 * cell index (the receiving slab's numbering).  Buffers are DEVICE memory, capacity particles each,
 * struct-of-arrays: member m at byte m*capacity*real_bytes, cell (int32) at byte 7*capacity*real_bytes.
 * cpic_append_particles_device appends n particles from such a buffer.  Both block the host. */
int  cpic_set_axis_periodic(cpic_ctx* ctx, int32_t px, int32_t py, int32_t pz);
/* The stencil kernels of advance_b / advance_e alone (src/fields.h:692-717 / 646-664, 534-543), without the
 * ghost fold / copy calls that the reference makes inside them (:718, :642-643): in slab mode the host
 * interleaves those with the neighbour exchange. */
int  cpic_advance_b_stencil(cpic_ctx* ctx, double px, double py, double pz);
int  cpic_advance_e_stencil(cpic_ctx* ctx, double px, double py, double pz, double dt_eps0);
int  cpic_extract_z_leavers(cpic_ctx* ctx, void* lo_buf, void* hi_buf, int64_t capacity, int64_t* n_lo, int64_t* n_hi,
                            int32_t rebase_lo, int32_t rebase_hi);
int  cpic_append_particles_device(cpic_ctx* ctx, const void* buf, int64_t capacity, int64_t n);
/* The same migration with every count kept ON THE DEVICE, so that the host never waits for the push and can
 * enqueue steps ahead (float reordering push only; CPIC_E_UNSUPPORTED otherwise).  extract: as
 * cpic_extract_z_leavers (leaver list of the last cpic_push_reorder), the two counts (n_lo, n_hi) are written
 * to counts_dev, int64[2] in DEVICE memory, for the caller to send along with the whole capacity-sized
 * buffers.  append: *count_dev (device) particles of a received buffer join the store.  The context's own
 * particle count lives on the device from the first extract on; any entry point that needs it on the host
 * (cpic_num_particles, downloads, cpic_push, cpic_sort_particles, ...) synchronises once and reports capacity
 * overflows that happened in between as CPIC_E_CAPACITY. */
int  cpic_slab_extract_async(cpic_ctx* ctx, void* lo_buf, void* hi_buf, int64_t capacity, int64_t* counts_dev,
                             int32_t rebase_lo, int32_t rebase_hi);
int  cpic_slab_append_async(cpic_ctx* ctx, const void* buf, int64_t capacity, const int64_t* count_dev);

/* Device-side timing of the last call of each kind, in milliseconds (CUDA events on the
 * context's stream).  what: 0 push, 1 sort, 2 field side (interp+unload+advance), 3 step total. */
int  cpic_last_ms(cpic_ctx* ctx, int what, double* ms);
int  cpic_launch_count(cpic_ctx* ctx, int64_t* launches);
/* Opt-in per-phase timing of cpic_step with CUDA events on the context's stream.  After a
 * profiled cpic_step, ms_out = total milliseconds over its steps spent in
 * [0] sort, [1] interpolator load + accumulator clear, [2] push (+mover+deposit), [3] field side. */
int  cpic_enable_step_profile(cpic_ctx* ctx, int32_t on);
int  cpic_step_profile(cpic_ctx* ctx, double ms_out[4], int64_t* steps);  /* kernels launched by this ctx so far */

#ifdef __cplusplus
}
#endif
#endif /* CABANAPIC_B200_H */
