/*
 * lbmk.h -- C ABI of a *generated* per-scheme kernel library (liblbmk_<hash>.so).
 *
 * One such library is produced by pylbm_b200.cudagen for every scheme dictionary
 * (nvcc, sm_100a) and replaces, for generator='cuda', the extension module that
 * the reference builds from its generated Cython source:
 *   - module attributes looked up by the driver: pylbm/algorithm/base.py:608-618,669-681
 *     (transport, f2m, m2f, relaxation, equilibrium, one_time_step[, source_term])
 *   - how they are called (kwargs by name): pylbm/symbolic.py:288-299
 *   - how the module is built/loaded: pylbm/generator/autowrap.py:52-139
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every launcher only
 * ENQUEUES on the given CUDA stream and returns 0, or a negative cudaError_t;
 * the caller owns every buffer; kernels never allocate.
 */
#ifndef LBMK_H
#define LBMK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBMK_ABI_VERSION 1

/*
 * Device array geometry.  Arrays are structure-of-arrays
 *     [population k][i0][i1][i2]       (canonical 3-D; 1-D and 2-D grids use n[0] (and n[1]) = 1)
 * with padded rows along the fastest axis i2:
 *     element(k, i0, i1, i2) = base + k*pstride + lead + (i0*n[1] + i1)*pitch + i2
 * `lead` is chosen so that the first interior cell of every row is 128-byte aligned
 * (replaces the reference's `sorder` permutation of a dense NumPy array,
 * pylbm/storage.py:60-118).  n[] INCLUDES the ghost layers (reference: vmax per side).
 * pitch, lead and pstride are counted in ELEMENTS and ONE grid describes both the input and the
 * output array of a launch, whatever their element types (fp32 populations in, fp64 moments out):
 * arrays that meet in a kernel are allocated with the same element layout.
 */
typedef struct {
    int n[3];         /* halo-inclusive logical sizes, slowest .. fastest            */
    int lo[3];        /* first index updated by the launch, per axis                 */
    int hi[3];        /* one past the last index updated by the launch, per axis     */
    int tx;           /* threads of a 128-thread block laid along i2 (power of two)  */
    int w[3];         /* ghost width per axis (interior = [w, n - w))                */
    int wrap;         /* bit a (a = 0..2) set: lbmk_one_time_step also stores, for the cells within
                         w of a face of axis a, their periodic images into the ghost layer
                         (= the ghost update of the NEXT step, storage.py:333-367);
                         LBMK_WRAP_PDL set: launch with programmatic stream serialization (the
                         kernel may be scheduled before the previous kernel of the stream has
                         finished; it touches memory only after griddepcontrol.wait)        */
    int64_t pitch;    /* elements between consecutive rows                           */
    int64_t lead;     /* position of logical index i2 = 0 inside a row               */
    int64_t pstride;  /* elements between consecutive populations                    */
} lbmk_grid;

#define LBMK_WRAP_PDL 0x100

/*
 * Neighbour slabs of a multi-GPU run (one process per GPU, x-slabs).  When given, the fused kernel
 * stores the populations leaving through its low / high slab face DIRECTLY into the ghost planes of
 * the neighbour rank's array (peer-mapped with CUDA IPC, NVLink) -- the per-step halo exchange of
 * the reference (mpi4py Isend/Irecv, storage.py:333-367) fused into the compute kernel.
 */
typedef struct {
    void* lo;            /* neighbour array receiving the images of my LOW-face cells (its high ghost) */
    void* hi;            /* neighbour array receiving the images of my HIGH-face cells (its low ghost) */
    int64_t pstride_lo;  /* population stride of those arrays                                          */
    int64_t pstride_hi;
    int nin_lo;          /* interior size of the `lo` neighbour along the slab axis                    */
} lbmk_peers;

/*
 * Bounce-back walls handled by the fused kernel itself.  When both faces normal to the FASTEST axis are
 * plain bounce-back / anti-bounce-back walls (reference: boundary.py:400-469, 626-684), every cell next
 * to such a wall stores, together with its new population moving towards the wall, the bounced value
 * f_sym(k)(c + v_k) = +-f_k(c) + rhs[k] into the wall's ghost cell: the boundary entries of the NEXT
 * step whose accesses are scattered 8-byte sectors in the list kernel (lbm_bc_apply), at the cost of
 * the periodic-image stores they replace.  The caller (boundary.plan_walls) proves that the entries
 * are complete, uniform per population and independent of every other entry.
 */
typedef struct {
    int lo_plane;      /* index along the fastest axis of the cells next to the LOW wall  */
    int hi_plane;      /* ... next to the HIGH wall                                       */
    int neg_lo;        /* 0: bounce-back (+f), 1: anti-bounce-back (-f)                   */
    int neg_hi;
    double rhs[64];    /* per LOADED population k (the one moving towards the wall)       */
} lbmk_walls;

/*
 * Boundary entries evaluated inside the fused kernel ("tasks").  The entry `f_k(c_out) = value` of a
 * boundary method (reference loops: boundary.py:462-464, 608-618, 678-680, 745-756, 818) is read by
 * exactly one pull of the following fused kernel: cell c_out + v_k, population k.  With a task table the
 * 128-thread block that owns the pulling cell evaluates its entries itself -- one entry per thread, the
 * arithmetic of the list kernel (lbm_bc_apply), values handed over through shared memory -- so a time
 * step is ONE launch and the scattered ghost stores / re-loads disappear.  The caller
 * (boundary.plan_tasks) proves that no entry reads a position another entry stores; then the input
 * array alone determines every value and the result is bit-identical to list kernels + pull.
 * All pointers are DEVICE pointers; tasks are sorted by block: block b (linear index
 * ((i0 - w[0]) * ngroups_y + row group) * ngroups_x + chunk of the fastest axis) owns
 * [block_ptr[b], block_ptr[b+1]).
 */
typedef struct {
    const int* block_ptr;          /* [nblocks + 1]                                               */
    const unsigned* code;          /* thread in block | population k << 8 | LBM_BC_* kind << 16    */
    const long long* l0;           /* element positions read in the INPUT array (lbmk_grid layout) */
    const long long* l1;           /* second load of the Bouzidi kinds                             */
    const double* const* rhs;      /* address of the entry's right-hand side (NULL: Neumann)       */
    const double* dist;            /* Bouzidi coefficient                                          */
    int ngroups_y, ngroups_x, tx;  /* launch geometry the table was built for (checked)            */
} lbmk_tasks;

/* ABI version the library was generated for. */
int lbmk_abi_version(void);

/* JSON description: {"abi", "dim", "nv", "storage", "routines": {name: {"scalars": [...],
 * "in", "out", "inner", "ops"}}}.  `scalars` gives, in order, the names of the runtime
 * scalars (t, dt, user parameters: pylbm/algorithm/base.py:630-667) expected in `scalars[]`. */
const char* lbmk_describe(void);

/*
 * Per-routine launcher; `fin`/`fout` are device pointers to arrays laid out as above.
 *   lbmk_one_time_step : fout = collide(pull-stream(fin)) on [lo, hi)   (pull.py:11-59)
 *   lbmk_transport     : fout(x) = fin(x - v_k)                         (base.py:289-296)
 *   lbmk_f2m / lbmk_m2f: m = M f / f = M^{-1} m on [lo, hi)             (base.py:336-379)
 *   lbmk_equilibrium, lbmk_relaxation, lbmk_source_term: in-place on m  (base.py:395-504)
 */
typedef int (*lbmk_launch_fn)(const void* fin, void* fout, const lbmk_grid* g,
                              const double* scalars, void* stream);

int lbmk_one_time_step(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
/* same with direct peer stores of the slab-face images (peers may be NULL) */
typedef int (*lbmk_launch_peers_fn)(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                    const lbmk_peers* peers, void* stream);
int lbmk_one_time_step_peers(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                             const lbmk_peers* peers, void* stream);
/* same with walls of the fastest axis applied by the kernel (walls may be NULL) */
typedef int (*lbmk_launch_walls_fn)(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                    const lbmk_peers* peers, const lbmk_walls* walls, void* stream);
int lbmk_one_time_step_walls(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                             const lbmk_peers* peers, const lbmk_walls* walls, void* stream);
/* same with the boundary entries of the step evaluated by the kernel (tasks may be NULL);
 * returns -4 when the table does not match the launch geometry */
typedef int (*lbmk_launch_tasks_fn)(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                                    const lbmk_peers* peers, const lbmk_tasks* tasks, void* stream);
int lbmk_one_time_step_tasks(const void* fin, void* fout, const lbmk_grid* g, const double* scalars,
                             const lbmk_peers* peers, const lbmk_tasks* tasks, void* stream);
/*
 * In-place streaming (AA pattern; only in a library generated for it).  ONE array.  phase 0, the even
 * step: cell x reads (k, x - v_k) like lbmk_one_time_step and stores its new population k into the slot it
 * read for the opposite population, (kbar, x + v_k) -- plus, for a population that leaves the interior, at
 * the fully wrapped position (its periodic image).  phase 1, the odd step: cell x reads (kbar, x), which IS
 * what a pull would bring, and stores (k, x) together with the periodic images of the next even step.
 * lbmk_f2m_sw / lbmk_f2m_consm_sw read the moments of an array that is in the swapped layout.
 */
typedef int (*lbmk_launch_aa_fn)(void* f, const lbmk_grid* g, const double* scalars, int phase, void* stream);
int lbmk_one_time_step_aa(void* f, const lbmk_grid* g, const double* scalars, int phase, void* stream);
typedef int (*lbmk_launch_aa_walls_fn)(void* f, const lbmk_grid* g, const double* scalars, int phase,
                                       const lbmk_walls* walls, void* stream);
/* in-place steps with the walls of the fastest axis applied by the kernel: the odd step stores the bounced
 * values into the wall's ghost cells like lbmk_one_time_step_walls, the even step into the cell's OWN slot
 * of the population moving towards the wall (where the following odd step reads what enters from the wall) */
int lbmk_one_time_step_aa_walls(void* f, const lbmk_grid* g, const double* scalars, int phase,
                                const lbmk_walls* walls, void* stream);
int lbmk_f2m_sw(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_f2m_consm_sw(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_transport(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_f2m(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
/* conserved moments only: fout has nconsm populations (rows 0..nconsm-1 of M f), same grid */
int lbmk_f2m_consm(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_m2f(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_equilibrium(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
int lbmk_relaxation(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);
/* only present when the scheme has source terms */
int lbmk_source_term(const void* fin, void* fout, const lbmk_grid* g, const double* scalars, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LBMK_H */
