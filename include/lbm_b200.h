/*
 * lbm_b200.h -- C ABI of the static B200 runtime (liblbm_b200.so).
 *
 * This is the drop-in boundary of the `generator='cuda'` backend: everything the
 * reference does per time step in Python/NumPy/mpi4py around its generated
 * kernels is behind these entry points, so that `Simulation.one_time_step()`
 * (pylbm/simulation.py:392-420) becomes ONE call that only enqueues work:
 *
 *   lbm_sim_step            <- Simulation.one_time_step          simulation.py:392-420
 *   lbm_sim_boundary_condition <- Simulation.boundary_condition  simulation.py:373-390
 *   lbm_periodic            <- Array.update (halo / periodic)    storage.py:306-367, 370-420
 *   lbm_sim_comm_init + slab exchange inside lbm_sim_step
 *                           <- mpi4py Irecv/Isend/Waitall        storage.py:210-303, 333-367
 *   lbm_bc_apply            <- generated bounce_back, Bouzidi_bounce_back, anti_bounce_back,
 *                              Bouzidi_anti_bounce_back, neumann* loops
 *                                                               boundary.py:462-464, 608-618,
 *                                                               678-680, 745-756, 818
 *   lbm_sim_add_bc / lbm_sim_set_rhs <- BoundaryMethod.move2gpu / set_rhs
 *                                                               boundary.py:378-397, 421-427
 *   lbm_sim_bc_groups, lbm_sim_set_walls, lbm_sim_bc_stale_only
 *                           <- the loop `for method in bc.methods: method.update(F)`
 *                              (simulation.py:387-390) with fewer launches: merged methods, and
 *                              plain bounce-back walls applied by the fused kernel itself
 *   lbm_array_h2d / lbm_array_d2h <- Array.__setitem__/__getitem__ device sync
 *                                                               storage.py:126-157
 *
 * Conventions: extern "C", plain pointers and sizes, no torch types.  Every
 * function returns 0 on success or a negative code (-cudaError_t, or -1000-ncclResult_t,
 * or -2000-x for argument errors); lbm_last_error() gives the message.  Nothing throws
 * or aborts.  The caller owns all buffers passed in; the runtime owns only what it
 * allocates itself (device copies of boundary lists, scratch, streams, events).
 * One host thread drives one GPU (one process per GPU); a context is not thread-safe (the staging
 * buffers of lbm_array_h2d/d2h are process-wide and unsynchronised: one copy at a time per process).
 *
 * Limits (violations return an argument error, nothing is truncated silently):
 *   populations per scheme          nv <= 64            (lbm_sim_desc.vel, lbmk_walls.rhs, PopSel)
 *   runtime scalars of a kernel     nscalars <= 32      (lbm_sim_desc.scalars: t, dt, user symbols)
 *   methods merged into one launch  <= 12               (lbm_sim_bc_groups; longer groups must be split
 *                                                        by the caller, see boundary.merge_groups)
 *   ghost width                     vmax[a] >= 0 per axis, n[a] >= 2*vmax[a] + 1
 *   peer halo rendezvous            a neighbour that is more than PYLBM_B200_HALO_TIMEOUT_S seconds
 *                                   late (default 30) makes the next lbm_sim_step / lbm_sim_sync /
 *                                   lbm_sim_timer_stop return -2003 "peer halo timeout"
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>
#include "lbmk.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_ABI_VERSION 1

#define LBM_STORAGE_F64 0
#define LBM_STORAGE_F32 1

/* boundary kernels (reference routine names in comments) */
#define LBM_BC_BOUNCE_BACK 0          /* f[s] =  f[l0] + rhs                 bounce_back              */
#define LBM_BC_ANTI_BOUNCE_BACK 1     /* f[s] = -f[l0] + rhs                 anti_bounce_back         */
#define LBM_BC_BOUZIDI_BOUNCE_BACK 2  /* f[s] =  d*c[l0] + (1-d)*c[l1] + rhs  Bouzidi_bounce_back (c = snapshot) */
#define LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK 3 /* f[s] = -d*f[l0] + (1-d)*f[l1] + rhs Bouzidi_anti_bounce_back */
#define LBM_BC_NEUMANN 4              /* f[s] =  f[l0]                       neumann, neumannx/y/z    */

int lbm_abi_version(void);
const char* lbm_last_error(void);

/* ---- device and memory -------------------------------------------------- */
int lbm_device_count(void);
int lbm_set_device(int device);
int lbm_device_sync(void);
int lbm_mem_info(uint64_t* free_bytes, uint64_t* total_bytes);
int lbm_malloc(void** ptr, uint64_t bytes);
int lbm_free(void* ptr);
int lbm_memset(void* ptr, int value, uint64_t bytes);
int lbm_host_alloc(void** ptr, uint64_t bytes);   /* pinned host memory */
int lbm_host_free(void* ptr);
int lbm_memcpy_h2d(void* dst, const void* src, uint64_t bytes);
int lbm_memcpy_d2h(void* dst, const void* src, uint64_t bytes);
int lbm_memcpy_d2d(void* dst, const void* src, uint64_t bytes);

/* Dense host block [nk][n0][n1][n2] of doubles <-> populations k0..k0+nk-1 of a padded
 * device array (see lbmk_grid).  Synchronous. */
int lbm_array_h2d(void* dev, const double* host, const lbmk_grid* g, int storage, int k0, int nk);
int lbm_array_d2h(double* host, const void* dev, const lbmk_grid* g, int storage, int k0, int nk);

/* ---- stand-alone operators (enqueue on `stream`, 0 = default stream) ------ */

/* Periodic ghost update of the axes selected by axis_mask (bit a = canonical axis a),
 * axes in increasing order so that edges/corners propagate like the reference's
 * dimension-by-dimension exchange.  vmax[a] = ghost width of axis a. */
int lbm_periodic(void* f, const lbmk_grid* g, int nv, int storage, const int vmax[3],
                 int axis_mask, void* stream);

/* One boundary kernel over ncond entries.  istore/iload0/iload1 are DEVICE arrays of
 * element positions (already including population and padding: see lbmk_grid), rhs/dist
 * DEVICE arrays of doubles (may be NULL when the kind does not use them).  two_phase != 0:
 * all loads are done before any store (scratch: DEVICE array of ncond doubles). */
int lbm_bc_apply(int kind, void* f, int storage, int64_t ncond, const int64_t* istore,
                 const int64_t* iload0, const int64_t* iload1, const double* rhs,
                 const double* dist, double* scratch, int two_phase, void* stream);

/* ---- time-step object ---------------------------------------------------- */
typedef struct lbm_sim lbm_sim;

typedef struct {
    int nv;                 /* number of populations Q                                   */
    int storage;            /* LBM_STORAGE_*                                             */
    lbmk_grid grid;         /* geometry; lo/hi = interior range of the fused kernel      */
    int vmax[3];            /* ghost width per canonical axis                            */
    int periodic_mask;      /* axes wrapped locally (bit a); axis 0 is exchanged between
                               ranks instead once lbm_sim_comm_init was called           */
    void* f;                /* device arrays (caller-owned)                              */
    void* fnew;
    lbmk_launch_fn one_time_step;  /* from the generated kernel library                  */
    lbmk_launch_peers_fn one_time_step_peers;   /* same, with peer stores (may be NULL)  */
    int nscalars;           /* runtime scalars of the fused kernel                       */
    int t_index;            /* position of `t` in scalars[], or -1                       */
    double scalars[32];
    double t;               /* current time, advanced by dt every step                   */
    double dt;
    int8_t vel[64][3];      /* lattice velocity of population k, canonical axes (slowest..fastest);
                               ghost layers are refreshed only for the populations that enter
                               the domain through them                                    */
} lbm_sim_desc;

lbm_sim* lbm_sim_create(const lbm_sim_desc* desc);
void lbm_sim_destroy(lbm_sim* sim);

/* Register one boundary method (order of calls = order of application, like
 * Simulation.bc.methods).  All arrays are HOST arrays and are copied to the device.
 * Entries are grouped in `nlevels` consecutive levels [level_ptr[i], level_ptr[i+1]);
 * levels run one after the other, entries of a level in parallel; two_phase[i] != 0
 * makes level i gather before it scatters.  Returns the index of the method (>= 0). */
int lbm_sim_add_bc(lbm_sim* sim, int kind, int64_t ncond, const int64_t* istore,
                   const int64_t* iload0, const int64_t* iload1, const double* rhs,
                   const double* dist, int nlevels, const int64_t* level_ptr,
                   const int* two_phase);
int lbm_sim_set_rhs(lbm_sim* sim, int ibc, const double* rhs_host);
/* Time-dependent boundary values without a host round trip (reference: boundary.py:307-321 update_feq +
 * 421-427 set_rhs, both NumPy on the host every step).  lbm_sim_upload_rows brings the moments the
 * user's callback wrote (small host block, `height` rows of `width` bytes) to the device on the
 * simulation's stream through page-locked staging, without stalling the host; the caller then enqueues
 * the equilibrium / m2f kernels of the kernel library on lbm_sim_stream() and lbm_sim_rhs_update, which
 * recomputes rhs[dst[j]] = feq[a[j]] + sign * feq[b[j]] (sign = -1 bounce-back kinds, +1 anti-bounce-back
 * kinds) for j < count in the method's DEVICE list.  a, b, dst, feq are DEVICE pointers. */
int lbm_sim_upload_rows(lbm_sim* sim, void* dst_dev, uint64_t dpitch, const void* src_host, uint64_t spitch,
                        uint64_t width, uint64_t height);
int lbm_sim_rhs_update(lbm_sim* sim, int ibc, int64_t count, const int64_t* a_dev, const int64_t* b_dev,
                       const int64_t* dst_dev, double sign, const double* feq_dev);
/* flag != 0: method `ibc` is applied only when the ghost layers of the current array are NOT the ones
 * the previous fused launch produced (first step, after lbm_sim_invalidate_ghosts).  Used for the
 * entries of the walls handed to lbm_sim_set_walls: in steady state the fused kernel has already
 * stored their values. */
int lbm_sim_bc_stale_only(lbm_sim* sim, int ibc, int flag);
/* Let the fused kernel apply the bounce-back walls of the fastest axis (lbmk_walls, lbmk.h);
 * `launcher` = lbmk_one_time_step_walls of the kernel library.  walls = NULL switches it off. */
int lbm_sim_set_walls(lbm_sim* sim, lbmk_launch_walls_fn launcher, const lbmk_walls* walls);
/* Let the fused kernel evaluate the boundary entries itself (lbmk_tasks, lbmk.h): a time step is then
 * ONE kernel launch.  `launcher` = lbmk_one_time_step_tasks of the kernel library.  HOST arrays, one
 * entry per task, sorted by block: code / l0 / l1 / dist as in lbmk_tasks; ibc[t] and entry[t] name
 * the registered method (lbm_sim_add_bc index) and the position inside its lists whose right-hand
 * side the task uses (so lbm_sim_set_rhs keeps working).  The registered methods stay: they are what
 * lbm_sim_boundary_condition applies.  ntasks < 0 switches the table off.  Not combined with
 * lbm_sim_set_walls. */
int lbm_sim_set_tasks(lbm_sim* sim, lbmk_launch_tasks_fn launcher, int64_t ntasks, int64_t nblocks,
                      const int32_t* block_ptr, const uint32_t* code, const int64_t* l0, const int64_t* l1,
                      const double* dist, const int32_t* ibc, const int64_t* entry,
                      int ngroups_y, int ngroups_x, int tx);
/* In-place streaming (AA pattern): the populations live in ONE array (desc.f == desc.fnew), half the
 * HBM footprint of the two-array pull scheme, same traffic per step.  `launcher` =
 * lbmk_one_time_step_aa of a kernel library generated for it.  Steps alternate between the even kernel
 * (natural -> swapped layout) and the odd kernel (swapped -> natural); after an even step every position
 * (k, y) of a boundary list is found at (kbar, y + v_k), so each method needs its lists once more in that
 * form (lbm_sim_set_bc_odd; boundary.plan_aa computes them).  On x-slabs (after lbm_sim_comm_init, NCCL
 * halo, not the peer halo) an even step is followed by a REVERSE exchange: the ghost planes, where the
 * even kernel deposited what left through the slab faces, travel to the neighbours' interior planes.
 * Not combined with lbm_sim_set_tasks; fused walls through lbm_sim_set_aa_walls.
 * lbm_sim_aa_phase: 0 natural, 1 swapped, -1 off. */
int lbm_sim_set_aa(lbm_sim* sim, lbmk_launch_aa_fn launcher);
int lbm_sim_set_bc_odd(lbm_sim* sim, int ibc, const int64_t* istore, const int64_t* iload0, const int64_t* iload1);
int lbm_sim_aa_phase(lbm_sim* sim);
/* in-place streaming with the bounce-back walls of the fastest axis applied by the kernels
 * (lbmk_one_time_step_aa_walls); the replaced entries stay registered as stale-only methods, with their
 * odd-step lists, exactly as for lbm_sim_set_walls.  After lbm_sim_set_aa. */
int lbm_sim_set_aa_walls(lbm_sim* sim, lbmk_launch_aa_walls_fn launcher, const lbmk_walls* walls);
/* Merged launches: the registered methods [group_ptr[g], group_ptr[g+1]) run as ONE kernel launch.
 * The caller must have proved that, inside a group, no entry reads or overwrites a position that an
 * entry of ANOTHER method of the group stores (results are then bit-identical to running the methods
 * one after the other; fewer launches matter for small lattices).  A method in a group of more than
 * one must be a single level without gather phase.  ngroups = 0 restores one launch sequence per
 * method; registering another method does the same. */
int lbm_sim_bc_groups(lbm_sim* sim, int ngroups, const int* group_ptr);
int lbm_sim_set_scalars(lbm_sim* sim, const double* scalars, int nscalars);

/* nsteps x (ghost update -> boundary methods -> fused pull stream+collide -> swap). */
int lbm_sim_step(lbm_sim* sim, int nsteps);
/* ghost update + boundary methods only, on the current f. */
int lbm_sim_boundary_condition(lbm_sim* sim);
int lbm_sim_sync(lbm_sim* sim);
/* current arrays (after the swaps), time and step count */
int lbm_sim_state(lbm_sim* sim, void** f, void** fnew, double* t, int64_t* nt);
int lbm_sim_set_state(lbm_sim* sim, void* f, void* fnew, double t);
/* The fused kernel writes the periodic images of the NEXT step's ghost update, so the copy
 * kernels are skipped after the first step.  Call this when F was modified from outside. */
int lbm_sim_invalidate_ghosts(lbm_sim* sim);
/* capture pairs of steps in a CUDA graph (only when the kernel does not depend on t) */
int lbm_sim_use_graph(lbm_sim* sim, int enable);
/* CUDA-event timer on the stream the kernels are launched on */
int lbm_sim_timer_start(lbm_sim* sim);
int lbm_sim_timer_stop(lbm_sim* sim, float* elapsed_ms);
/* per-launch timing of the fused kernel: while enabled every fused-kernel launch is bracketed
 * by two CUDA events on the launch stream (CUDA graphs are bypassed); profile_read returns the
 * summed device time and the number of launches since the last read. */
int lbm_sim_profile(lbm_sim* sim, int enable);
int lbm_sim_profile_read(lbm_sim* sim, double* fused_ms, int64_t* nlaunch);
/* number of kernels launched by this object so far */
int64_t lbm_sim_launch_count(lbm_sim* sim);
void* lbm_sim_stream(lbm_sim* sim);

/* ---- multi-GPU: x-slabs, one process per GPU, NCCL send/recv over NVLink ---- */
int lbm_comm_unique_id(void* id128);   /* 128-byte ncclUniqueId, created on rank 0 */
int lbm_sim_comm_init(lbm_sim* sim, int rank, int nranks, const void* id128);

/* Direct NVLink halo: every rank exports CUDA-IPC handles of its two population arrays and of a
 * small flag buffer (lbm_sim_ipc_export: 3 x 64 bytes + geometry), gathers the blobs of its two ring
 * neighbours by any means (torch.distributed, a queue, ...) and opens them (lbm_sim_ipc_open).
 * From then on the fused kernel stores the slab-face images straight into the neighbours' ghost
 * planes and the ranks synchronise with device-side arrival counters: no NCCL call per step. */
#define LBM_IPC_BLOB_BYTES 256
int lbm_sim_ipc_export(lbm_sim* sim, void* blob256);
int lbm_sim_ipc_open(lbm_sim* sim, const void* left_blob256, const void* right_blob256);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
