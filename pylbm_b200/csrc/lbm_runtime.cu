// lbm_runtime.cu -- static B200 runtime behind include/lbm_b200.h (sm_100a).
//
// What the reference does per step in Python around its generated kernels
// (pylbm/simulation.py:373-420) lives here as one enqueue-only C++ driver:
//   periodic / slab ghost update  (pylbm/storage.py:306-367)
//   boundary kernels              (pylbm/boundary.py:462-464, 608-618, 678-680, 745-756, 818)
//   fused pull stream+collide     (generated library, include/lbmk.h)
//   F <-> Fnew swap               (pylbm/simulation.py:417)
// All kernels here are HBM/latency-bound gather/scatter or plane copies: no tensor cores.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "lbm_b200.h"

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_error(int code, const char* what, const char* detail) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail ? detail : "");
    return code;
}

#define CUDA_TRY(call)                                                            \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return set_error(-(int)e_, #call, cudaGetErrorString(e_)); \
    } while (0)

#define ARG_ERROR(msg) set_error(-2001, "argument error", msg)

// ---------------------------------------------------------------------------
// programmatic dependent launch (sm_90+): the kernels of a time step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization.  Each of them starts with
//     griddepcontrol.launch_dependents;   the next kernel of the stream may be scheduled as soon as all
//                                         blocks of this one have started
//     griddepcontrol.wait;                nothing is read or written before the previous kernel has
//                                         completed and its memory operations are visible
// so results are unchanged and the kernel-to-kernel latency of a step (1-2 us per dependent launch,
// most of the 5 us step of a 256^2 lattice) overlaps with the tail of the previous kernel.  Both
// instructions are no-ops in a kernel launched without the attribute.
// ---------------------------------------------------------------------------
#define LBM_PDL_PROLOGUE()                                         \
    asm volatile("griddepcontrol.launch_dependents;");             \
    asm volatile("griddepcontrol.wait;" ::: "memory")

template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static bool g_pdl = true;      // PYLBM_B200_NO_PDL=1 switches it off (lbm_sim_create reads the variable)

extern "C" int lbm_abi_version(void) { return LBM_ABI_VERSION; }
extern "C" const char* lbm_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------
// device and memory
// ---------------------------------------------------------------------------
extern "C" int lbm_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return set_error(-(int)e, "cudaGetDeviceCount", cudaGetErrorString(e));
    return n;
}
// Page-locked host buffers (lbm_host_alloc, the staging side of lbm_array_h2d/d2h) should live on the
// NUMA node the GPU hangs off: with one process per GPU on a two-socket host, half of the ranks
// otherwise DMA across the socket interconnect.  Sets the calling thread's memory policy to "prefer
// the GPU's node" (threads created later inherit it); CPU affinity is left alone.  Best effort:
// any failure leaves the default policy.  PYLBM_B200_NO_NUMA=1 disables it.
static void prefer_gpu_numa_node(int device) {
    if (getenv("PYLBM_B200_NO_NUMA")) return;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* fh = fopen(path, "r");
    if (!fh) return;
    int node = -1;
    if (fscanf(fh, "%d", &node) != 1) node = -1;
    fclose(fh);
    if (node < 0 || node >= 1024) return;
    unsigned long mask[1024 / (8 * sizeof(unsigned long))];
    memset(mask, 0, sizeof(mask));
    mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
    const int MPOL_PREFERRED_ = 1;
    syscall(SYS_set_mempolicy, MPOL_PREFERRED_, mask, (unsigned long)(1024 + 1));
}

extern "C" int lbm_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    prefer_gpu_numa_node(device);
    return 0;
}
extern "C" int lbm_device_sync(void) { CUDA_TRY(cudaDeviceSynchronize()); return 0; }
extern "C" int lbm_mem_info(uint64_t* free_bytes, uint64_t* total_bytes) {
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return 0;
}
extern "C" int lbm_malloc(void** ptr, uint64_t bytes) {
    if (!ptr) return ARG_ERROR("null ptr");
    CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 8));
    return 0;
}
extern "C" int lbm_free(void* ptr) { CUDA_TRY(cudaFree(ptr)); return 0; }
extern "C" int lbm_memset(void* ptr, int value, uint64_t bytes) { CUDA_TRY(cudaMemset(ptr, value, bytes)); return 0; }
extern "C" int lbm_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return ARG_ERROR("null ptr");
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 8));
    return 0;
}
extern "C" int lbm_host_free(void* ptr) { CUDA_TRY(cudaFreeHost(ptr)); return 0; }
extern "C" int lbm_memcpy_h2d(void* dst, const void* src, uint64_t bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int lbm_memcpy_d2h(void* dst, const void* src, uint64_t bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int lbm_memcpy_d2d(void* dst, const void* src, uint64_t bytes) {
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    return 0;
}

// ---------------------------------------------------------------------------
// dense host block <-> padded device array
// ---------------------------------------------------------------------------
template <typename S, bool TO_PADDED>
__global__ void k_repack(S* __restrict__ padded, double* __restrict__ dense, lbmk_grid g, int k0, long long first,
                         long long count) {
    // dense: [nk][n0][n1][n2] flattened; this launch handles the dense elements [first, first + count),
    // staged in dense[0 .. count)
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const long long i = first + j;
    const long long n2 = g.n[2], n1 = g.n[1];
    const long long per_pop = (long long)g.n[0] * n1 * n2;
    const long long k = i / per_pop;
    const long long r = i - k * per_pop;
    const long long i2 = r % n2;
    const long long row = r / n2;  // i0*n1 + i1
    const long long pos = (k0 + k) * g.pstride + g.lead + row * g.pitch + i2;
    if (TO_PADDED)
        padded[pos] = (S)dense[j];
    else
        dense[j] = (double)padded[pos];
}

// two staging buffers + two streams per process: the repack kernel of one chunk overlaps the PCIe copy
// of the other (full speed when the host side is page-locked, e.g. lbm_host_alloc memory)
struct RepackPipe {
    int device = -1;
    double* stage[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    long long chunk = 0;   // elements per staging buffer
};
static RepackPipe g_pipe;

static int repack_pipe(long long want) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (g_pipe.device == dev && g_pipe.chunk >= want) return 0;
    for (int b = 0; b < 2; ++b) {
        if (g_pipe.stage[b]) cudaFree(g_pipe.stage[b]);
        if (g_pipe.st[b]) cudaStreamDestroy(g_pipe.st[b]);
        g_pipe.stage[b] = nullptr;
        g_pipe.st[b] = nullptr;
    }
    g_pipe.device = -1;
    if (!g_pipe.ready) CUDA_TRY(cudaEventCreateWithFlags(&g_pipe.ready, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaMalloc(&g_pipe.stage[b], (size_t)want * 8));
        CUDA_TRY(cudaStreamCreateWithFlags(&g_pipe.st[b], cudaStreamNonBlocking));
    }
    g_pipe.device = dev;
    g_pipe.chunk = want;
    return 0;
}

template <bool TO_PADDED>
static int repack(void* dev, double* host, const lbmk_grid* g, int storage, int k0, int nk) {
    if (!dev || !host || !g) return ARG_ERROR("null pointer");
    const long long per_pop = (long long)g->n[0] * g->n[1] * g->n[2];
    const long long total = per_pop * nk;
    if (total <= 0) return 0;
    const long long chunk = total < (8LL << 20) ? total : (8LL << 20);     // 64 MB of doubles per buffer
    int rc = repack_pipe(chunk);
    if (rc) return rc;
    // everything enqueued so far on the blocking streams (the simulations' streams) precedes the legacy
    // default stream; the two copy streams wait for that point on the device, the host does not stall
    CUDA_TRY(cudaEventRecord(g_pipe.ready, 0));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamWaitEvent(g_pipe.st[i], g_pipe.ready, 0));
    cudaError_t e = cudaSuccess;
    int b = 0;
    for (long long first = 0; first < total && e == cudaSuccess; first += chunk, b ^= 1) {
        const long long count = (total - first < chunk) ? total - first : chunk;
        const unsigned blocks = (unsigned)((count + 255) / 256);
        cudaStream_t st = g_pipe.st[b];
        double* stage = g_pipe.stage[b];
        if (TO_PADDED) {
            e = cudaMemcpyAsync(stage, host + first, (size_t)count * 8, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) break;
            if (storage == LBM_STORAGE_F64)
                k_repack<double, true><<<blocks, 256, 0, st>>>((double*)dev, stage, *g, k0, first, count);
            else
                k_repack<float, true><<<blocks, 256, 0, st>>>((float*)dev, stage, *g, k0, first, count);
            e = cudaGetLastError();
        } else {
            if (storage == LBM_STORAGE_F64)
                k_repack<double, false><<<blocks, 256, 0, st>>>((double*)dev, stage, *g, k0, first, count);
            else
                k_repack<float, false><<<blocks, 256, 0, st>>>((float*)dev, stage, *g, k0, first, count);
            e = cudaGetLastError();
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(host + first, stage, (size_t)count * 8, cudaMemcpyDeviceToHost, st);
        }
    }
    for (int i = 0; i < 2; ++i) {
        cudaError_t e2 = cudaStreamSynchronize(g_pipe.st[i]);
        if (e == cudaSuccess) e = e2;
    }
    if (e != cudaSuccess) return set_error(-(int)e, "lbm_array copy", cudaGetErrorString(e));
    return 0;
}

extern "C" int lbm_array_h2d(void* dev, const double* host, const lbmk_grid* g, int storage, int k0, int nk) {
    return repack<true>(dev, const_cast<double*>(host), g, storage, k0, nk);
}
extern "C" int lbm_array_d2h(double* host, const void* dev, const lbmk_grid* g, int storage, int k0, int nk) {
    return repack<false>(const_cast<void*>(dev), host, g, storage, k0, nk);
}

// ---------------------------------------------------------------------------
// periodic ghost update (reference: storage.py:333-367 with one rank, 370-420 on GPU)
// ---------------------------------------------------------------------------
// For axis A with ghost width w:  ghost [0,w) <- [n-2w, n-w),  ghost [n-w, n) <- [w, 2w)
// over the FULL extent of the other axes (ghosts included), so that doing the axes in increasing
// order fills edges and corners like the reference.  `sel` lists the populations to copy and, for
// each, which ghost side: the low ghost layer is only ever read for populations moving in +A, the
// high one for populations moving in -A (see cudagen.py, periodic images), so the time-step driver
// passes the sign-matched list; lbm_periodic() passes every population and both sides.
// The extent of axis 0 can be restricted to [x0, x1).
struct PopSel {
    int n;
    unsigned char k[64];
    unsigned char side[64];   // bit 0: low ghost, bit 1: high ghost
};

template <typename S, int A>
__global__ void k_periodic(S* __restrict__ f, lbmk_grid g, PopSel sel, int w, int x0, int x1) {
    const long long n2 = g.n[2], n1 = g.n[1];
    // iteration space (population entry, e0, e1, e2) where axis A has extent w
    const long long e0 = (A == 0) ? w : (x1 - x0);
    const long long e1 = (A == 1) ? w : n1;
    const long long e2 = (A == 2) ? w : n2;
    const long long total = (long long)sel.n * e0 * e1 * e2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const long long c2 = r % e2; r /= e2;
        const long long c1 = r % e1; r /= e1;
        const long long c0 = r % e0; r /= e0;
        const int p = (int)r;
        // coordinates with the one along A left at 0 (j runs over the ghost layers of A)
        const long long i0 = (A == 0) ? 0 : x0 + c0, i1 = (A == 1) ? 0 : c1, i2 = (A == 2) ? 0 : c2;
        const long long j = (A == 0) ? c0 : (A == 1) ? c1 : c2;
        const long long n = g.n[A];
        const long long stride = (A == 0) ? n1 * g.pitch : (A == 1) ? g.pitch : 1;
        S* base = f + ((long long)sel.k[p] * g.pstride + g.lead + (i0 * n1 + i1) * g.pitch + i2);
        if (sel.side[p] & 1) base[j * stride] = base[(n - 2 * w + j) * stride];
        if (sel.side[p] & 2) base[(n - w + j) * stride] = base[(w + j) * stride];
    }
}

template <typename S>
static cudaError_t launch_periodic(S* f, const lbmk_grid& g, const PopSel& sel, const int vmax[3], int axis, int x0,
                                   int x1, cudaStream_t st, int64_t* nlaunch) {
    const int w = vmax[axis];
    if (w <= 0 || g.n[axis] < 2 * w + 1 || sel.n == 0) return cudaSuccess;
    if (axis != 0 && x1 <= x0) return cudaSuccess;
    long long e0 = (axis == 0) ? w : (x1 - x0);
    long long e1 = (axis == 1) ? w : g.n[1];
    long long e2 = (axis == 2) ? w : g.n[2];
    long long total = (long long)sel.n * e0 * e1 * e2;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    if (blocks < 1) blocks = 1;
    if (axis == 0) k_periodic<S, 0><<<(unsigned)blocks, 256, 0, st>>>(f, g, sel, w, x0, x1);
    if (axis == 1) k_periodic<S, 1><<<(unsigned)blocks, 256, 0, st>>>(f, g, sel, w, x0, x1);
    if (axis == 2) k_periodic<S, 2><<<(unsigned)blocks, 256, 0, st>>>(f, g, sel, w, x0, x1);
    if (nlaunch) ++*nlaunch;
    return cudaGetLastError();
}

// sel[a] = selection for axis a
static cudaError_t periodic_axes(void* f, const lbmk_grid& g, const PopSel sel[3], int storage, const int vmax[3],
                                 int mask, int x0, int x1, cudaStream_t st, int64_t* nlaunch) {
    for (int a = 0; a < 3; ++a) {
        if (!(mask & (1 << a))) continue;
        cudaError_t e = (storage == LBM_STORAGE_F64)
                            ? launch_periodic<double>((double*)f, g, sel[a], vmax, a, x0, x1, st, nlaunch)
                            : launch_periodic<float>((float*)f, g, sel[a], vmax, a, x0, x1, st, nlaunch);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

extern "C" int lbm_periodic(void* f, const lbmk_grid* g, int nv, int storage, const int vmax[3], int axis_mask,
                            void* stream) {
    if (!f || !g || !vmax) return ARG_ERROR("null pointer");
    if (nv < 1 || nv > 64) return ARG_ERROR("nv must be in [1, 64]");
    PopSel sel[3];
    for (int a = 0; a < 3; ++a) {
        sel[a].n = nv;
        for (int k = 0; k < nv; ++k) { sel[a].k[k] = (unsigned char)k; sel[a].side[k] = 3; }
    }
    cudaError_t e = periodic_axes(f, *g, sel, storage, vmax, axis_mask, 0, g->n[0], (cudaStream_t)stream, nullptr);
    if (e != cudaSuccess) return set_error(-(int)e, "lbm_periodic", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------
// boundary kernels
// ---------------------------------------------------------------------------
// Arithmetic is written with explicit round-to-nearest intrinsics in the association
// order of the reference's generated C (no FMA contraction), so fp64 results are
// bit-identical to the sequential reference loop as long as entries are independent
// (dependent entries are separated into levels by the host, see boundary.py).
template <int KIND>
__device__ __forceinline__ double bc_value(double a, double b, double rhs, double d) {
    if (KIND == LBM_BC_BOUNCE_BACK) return __dadd_rn(a, rhs);
    if (KIND == LBM_BC_ANTI_BOUNCE_BACK) return __dadd_rn(-a, rhs);
    if (KIND == LBM_BC_BOUZIDI_BOUNCE_BACK)
        return __dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, d), b), __dmul_rn(d, a)), rhs);
    if (KIND == LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK)
        return __dadd_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, d), b), -__dmul_rn(d, a)), rhs);
    return a;  // Neumann
}

// PHASE 0: load + store in place; PHASE 1: load -> scratch; PHASE 2: scratch -> store
template <typename S, int KIND, int PHASE>
__global__ void k_bc(S* __restrict__ f, long long ncond, const long long* __restrict__ istore,
                     const long long* __restrict__ iload0, const long long* __restrict__ iload1,
                     const double* __restrict__ rhs, const double* __restrict__ dist, double* __restrict__ scratch) {
    LBM_PDL_PROLOGUE();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncond) return;
    if (PHASE == 2) {
        f[istore[i]] = (S)scratch[i];
        return;
    }
    constexpr bool two_loads = (KIND == LBM_BC_BOUZIDI_BOUNCE_BACK || KIND == LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK);
    const double a = (double)f[iload0[i]];
    const double b = two_loads ? (double)f[iload1[i]] : 0.0;
    const double r = (KIND == LBM_BC_NEUMANN) ? 0.0 : rhs[i];
    const double d = two_loads ? dist[i] : 0.0;
    const double v = bc_value<KIND>(a, b, r, d);
    if (PHASE == 0)
        f[istore[i]] = (S)v;
    else
        scratch[i] = v;
}

template <typename S, int KIND>
static cudaError_t launch_bc_kind(S* f, long long n, const long long* is, const long long* l0, const long long* l1,
                                  const double* rhs, const double* dist, double* scratch, int two_phase,
                                  cudaStream_t st, int64_t* nlaunch) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (!two_phase) {
        cudaError_t e = launch_k(k_bc<S, KIND, 0>, dim3(blocks), dim3(128), st, g_pdl, f, n, is, l0, l1, rhs, dist, scratch);
        if (nlaunch) ++*nlaunch;
        return e;
    }
    cudaError_t e = launch_k(k_bc<S, KIND, 1>, dim3(blocks), dim3(128), st, g_pdl, f, n, is, l0, l1, rhs, dist, scratch);
    if (e == cudaSuccess)
        e = launch_k(k_bc<S, KIND, 2>, dim3(blocks), dim3(128), st, g_pdl, f, n, is, l0, l1, rhs, dist, scratch);
    if (nlaunch) *nlaunch += 2;
    return e;
}

template <typename S>
static cudaError_t launch_bc(int kind, S* f, long long n, const long long* is, const long long* l0,
                             const long long* l1, const double* rhs, const double* dist, double* scratch,
                             int two_phase, cudaStream_t st, int64_t* nlaunch) {
    switch (kind) {
        case LBM_BC_BOUNCE_BACK:
            return launch_bc_kind<S, LBM_BC_BOUNCE_BACK>(f, n, is, l0, l1, rhs, dist, scratch, two_phase, st, nlaunch);
        case LBM_BC_ANTI_BOUNCE_BACK:
            return launch_bc_kind<S, LBM_BC_ANTI_BOUNCE_BACK>(f, n, is, l0, l1, rhs, dist, scratch, two_phase, st, nlaunch);
        case LBM_BC_BOUZIDI_BOUNCE_BACK:
            return launch_bc_kind<S, LBM_BC_BOUZIDI_BOUNCE_BACK>(f, n, is, l0, l1, rhs, dist, scratch, two_phase, st, nlaunch);
        case LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK:
            return launch_bc_kind<S, LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK>(f, n, is, l0, l1, rhs, dist, scratch, two_phase, st, nlaunch);
        case LBM_BC_NEUMANN:
            return launch_bc_kind<S, LBM_BC_NEUMANN>(f, n, is, l0, l1, rhs, dist, scratch, two_phase, st, nlaunch);
    }
    return cudaErrorInvalidValue;
}

// All methods of a step in ONE launch, for the common case where the host has proved that no entry
// of any method reads or overwrites what another entry stores (then the order of the methods and of
// their entries is irrelevant).  Same arithmetic as k_bc (bit-identical results).
#define LBM_BC_MAX_SEGMENTS 12
struct BcSegment {
    long long begin;   // first global entry index of this method
    int kind;
    const long long *istore, *iload0, *iload1;
    const double *rhs, *dist;
};
struct BcSegments {
    int n;
    long long total;
    BcSegment seg[LBM_BC_MAX_SEGMENTS];
};

template <typename S>
__global__ void k_bc_multi(S* __restrict__ f, const BcSegments segs) {
    LBM_PDL_PROLOGUE();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= segs.total) return;
    int m = segs.n - 1;
    while (m > 0 && i < segs.seg[m].begin) --m;
    const BcSegment& sg = segs.seg[m];
    const long long j = i - sg.begin;
    const double a = (double)f[sg.iload0[j]];
    double v;
    switch (sg.kind) {
        case LBM_BC_BOUNCE_BACK: v = bc_value<LBM_BC_BOUNCE_BACK>(a, 0.0, sg.rhs[j], 0.0); break;
        case LBM_BC_ANTI_BOUNCE_BACK: v = bc_value<LBM_BC_ANTI_BOUNCE_BACK>(a, 0.0, sg.rhs[j], 0.0); break;
        case LBM_BC_BOUZIDI_BOUNCE_BACK:
            v = bc_value<LBM_BC_BOUZIDI_BOUNCE_BACK>(a, (double)f[sg.iload1[j]], sg.rhs[j], sg.dist[j]);
            break;
        case LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK:
            v = bc_value<LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK>(a, (double)f[sg.iload1[j]], sg.rhs[j], sg.dist[j]);
            break;
        default: v = a; break;   // Neumann
    }
    f[sg.istore[j]] = (S)v;
}

extern "C" int lbm_bc_apply(int kind, void* f, int storage, int64_t ncond, const int64_t* istore,
                            const int64_t* iload0, const int64_t* iload1, const double* rhs, const double* dist,
                            double* scratch, int two_phase, void* stream) {
    if (!f || (ncond > 0 && (!istore || !iload0))) return ARG_ERROR("null pointer");
    // stand-alone call: Bouzidi bounce-back always reads a snapshot (the reference's fcopy)
    if (kind == LBM_BC_BOUZIDI_BOUNCE_BACK) two_phase = 1;
    if (two_phase && ncond > 0 && !scratch)
        return ARG_ERROR("two-phase boundary kernel needs a scratch buffer");
    cudaError_t e =
        (storage == LBM_STORAGE_F64)
            ? launch_bc<double>(kind, (double*)f, ncond, (const long long*)istore, (const long long*)iload0,
                                (const long long*)iload1, rhs, dist, scratch, two_phase, (cudaStream_t)stream, nullptr)
            : launch_bc<float>(kind, (float*)f, ncond, (const long long*)istore, (const long long*)iload0,
                               (const long long*)iload1, rhs, dist, scratch, two_phase, (cudaStream_t)stream, nullptr);
    if (e != cudaSuccess) return set_error(-(int)e, "lbm_bc_apply", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency; shares the copy torch already loaded)
// ---------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8;

static int nccl_load() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return set_error(-1099, "dlopen(libnccl.so.2)", dlerror());
#define NCCL_SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                      \
    if (!g_nccl.field) return set_error(-1098, "dlsym", name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return 0;
}

struct CommEntry { nccl_comm comm; int rank, nranks; };
static std::map<std::string, CommEntry> g_comms;

#define NCCL_TRY(call)                                                                  \
    do {                                                                                \
        int r_ = (call);                                                                \
        if (r_ != 0) return set_error(-1000 - r_, #call, g_nccl.GetErrorString(r_));    \
    } while (0)

extern "C" int lbm_comm_unique_id(void* id128) {
    if (!id128) return ARG_ERROR("null id");
    int rc = nccl_load();
    if (rc) return rc;
    nccl_uid id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}

// ---------------------------------------------------------------------------
// device-side synchronisation between neighbour ranks (direct NVLink halo)
// ---------------------------------------------------------------------------
// After its fused kernel of step s a rank increments an arrival counter in each neighbour's memory;
// before touching the ghost planes of step s+1 it waits until both neighbours have arrived s+1 times.
// Counters only grow and the expected value lives on the device, so the same two kernels can be
// replayed from a CUDA graph.
__global__ void k_signal(unsigned long long* to_left, unsigned long long* to_right) {
    LBM_PDL_PROLOGUE();
    __threadfence_system();
    if (threadIdx.x == 0) atomicAdd_system(to_left, 1ULL);    // I am the RIGHT neighbour of my left rank
    if (threadIdx.x == 1) atomicAdd_system(to_right, 1ULL);   // and the LEFT neighbour of my right rank
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// `err` is a word of page-locked host memory mapped into the device: a neighbour that does not arrive
// within `timeout_ns` (a dead or hung rank) makes the wait give up, record which side was late
// (bit 0: left, bit 1: right) and return, so that the step sequence drains instead of spinning for
// ever; the host reports it at the next lbm_sim_sync / lbm_sim_timer_stop / lbm_sim_step.  Once the
// word is set every later wait returns at once (the results are void from the first miss on).
__global__ void k_wait(unsigned long long* flags, volatile unsigned int* err, unsigned long long timeout_ns) {
    LBM_PDL_PROLOGUE();
    // flags[0], flags[1]: arrivals from the left / right neighbour; flags[2], flags[3]: my epochs
    const int i = threadIdx.x;
    if (i < 2) {
        const unsigned long long want = flags[2 + i] + 1ULL;
        volatile unsigned long long* arr = flags + i;
        if (*arr < want && *err == 0u) {
            const unsigned long long t0 = global_ns();
            unsigned int spins = 0;
            while (*arr < want) {
                __nanosleep(100);
                if ((++spins & 1023u) == 0u && global_ns() - t0 > timeout_ns) {
                    atomicOr_system((unsigned int*)err, 1u << i);
                    break;
                }
            }
        }
        flags[2 + i] = want;
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------
// time-step object
// ---------------------------------------------------------------------------
struct BcMethod {
    int kind = 0;
    int stale_only = 0;      // skipped when the ghost layers come from the previous fused launch
    long long ncond = 0;
    long long *istore = nullptr, *iload0 = nullptr, *iload1 = nullptr;
    // the same positions seen through an in-place array after an even step: (k, y) -> (kbar, y + v_k)
    long long *istore_odd = nullptr, *iload0_odd = nullptr, *iload1_odd = nullptr;
    double *rhs = nullptr, *dist = nullptr;
    std::vector<long long> level_ptr;
    std::vector<int> two_phase;
};

struct lbm_sim {
    lbm_sim_desc d;
    void *f = nullptr, *fnew = nullptr;
    double t = 0.0;
    int64_t nt = 0;
    std::vector<BcMethod> bcs;
    std::vector<int> bc_groups;                   // first method of each merged launch (+ end); empty: none
    lbmk_launch_walls_fn walls_fn = nullptr;      // fused kernel applies the walls of the fastest axis
    lbmk_walls walls;
    lbmk_launch_aa_fn aa_fn = nullptr;            // in-place streaming: ONE array, even / odd steps
    int aa_swapped = 0;                           // the array is in the swapped layout (after an even step)
    lbmk_launch_aa_walls_fn aa_walls_fn = nullptr;   // in-place steps that also apply the fused walls
    int aa_even_walled = 0;                       // the last even step stored the wall values itself
    lbmk_launch_tasks_fn tasks_fn = nullptr;      // fused kernel evaluates the boundary entries itself
    lbmk_tasks tasks;                             // device arrays owned by this object
    double* scratch = nullptr;
    long long scratch_n = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_ready = nullptr, ev_comm = nullptr;
    int64_t launches = 0;
    // slab exchange
    nccl_comm comm = nullptr;
    int rank = 0, nranks = 1;
    int slab_axis = 0;
    // direct NVLink halo (CUDA IPC): peer arrays [side lo/hi][buffer A/B], arrival counters
    void* up_buf[2] = {nullptr, nullptr};         // page-locked staging of lbm_sim_upload_rows
    uint64_t up_bytes[2] = {0, 0};
    cudaEvent_t up_event[2] = {nullptr, nullptr};
    int up_next = 0;
    void* xbuf[4] = {nullptr, nullptr, nullptr, nullptr};   // NCCL halo: packed send / receive planes
    size_t xbuf_bytes[4] = {0, 0, 0, 0};
    int xchg_inflight = 0;                        // NCCL halo: the exchange for the CURRENT f was issued on
                                                  // comm_stream during the previous step (ev_comm marks its end)
    int overlap = 1;                              // NCCL halo: exchange of step s+1 || inner cells of step s
    int peers_ready = 0;
    long long signals = 0, waits = 0;             // enqueued so far (host-side bookkeeping)
    int waited = 0;                               // the wait for the current f was already enqueued
    void* buf[2] = {nullptr, nullptr};            // my arrays A (= desc.f) and B (= desc.fnew)
    void* peer_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    long long peer_pstride[2] = {0, 0};
    int peer_nin_lo = 0;
    unsigned long long* flags = nullptr;          // [0],[1]: arrivals from left/right; [2],[3]: my epochs
    unsigned long long* peer_flags[2] = {nullptr, nullptr};
    unsigned int* wait_err = nullptr;             // page-locked, mapped: set by k_wait on a timeout
    unsigned int* wait_err_dev = nullptr;         // device alias of wait_err
    unsigned long long wait_timeout_ns = 30ULL * 1000000000ULL;
    // CUDA graph of two consecutive steps (f->fnew, fnew->f)
    int use_graph = 0;
    cudaGraphExec_t graph = nullptr;
    void* graph_f = nullptr;
    long long graph_signals = 0;
    int64_t graph_launches = 0;
    // optional per-launch timing of the fused kernel (CUDA events on the launch stream)
    // ghost layers of `f` already hold the periodic images (written by the previous fused launch)
    int ghost_fresh = 0;
    int wrap_mask = 0;       // axes whose images the fused kernel writes
    PopSel sel[3];           // sign-matched populations per axis for the copy kernels / exchange
    int profile = 0;
    std::vector<cudaEvent_t> prof_events;   // pairs (before, after)
    size_t prof_used = 0;
};

extern "C" lbm_sim* lbm_sim_create(const lbm_sim_desc* desc) {
    if (!desc || !desc->f || !desc->fnew || !desc->one_time_step) {
        ARG_ERROR("lbm_sim_create: null descriptor / arrays / kernel");
        return nullptr;
    }
    if (desc->nv < 1 || desc->nv > 64 || desc->nscalars < 0 || desc->nscalars > 32) {
        ARG_ERROR("lbm_sim_create: nv must be in [1,64] and nscalars in [0,32]");
        return nullptr;
    }
    lbm_sim* s = new lbm_sim();
    s->d = *desc;
    s->f = desc->f;
    s->fnew = desc->fnew;
    s->buf[0] = desc->f;
    s->buf[1] = desc->fnew;
    s->t = desc->t;
    // slab axis = first real axis of the canonical 3-D grid
    s->slab_axis = 0;
    while (s->slab_axis < 2 && desc->grid.n[s->slab_axis] == 1 && desc->vmax[s->slab_axis] == 0) ++s->slab_axis;
    // Blocking streams: they order themselves against the legacy default stream, on which the
    // synchronous host<->device array copies and memsets of the C ABI run.
    for (int a = 0; a < 3; ++a) {
        const int w = desc->vmax[a];
        if ((desc->periodic_mask & (1 << a)) && w > 0 && desc->grid.n[a] - 2 * w >= 2 * w) s->wrap_mask |= (1 << a);
    }
    g_pdl = getenv("PYLBM_B200_NO_PDL") == nullptr;
    if (getenv("PYLBM_B200_NO_ZWRAP")) s->wrap_mask &= ~(1 << 2);   // debugging aid: lean copy kernel for z
    if (getenv("PYLBM_B200_NO_WRAP")) s->wrap_mask = 0;             // debugging aid: copy kernels every step
    for (int a = 0; a < 3; ++a) {
        s->sel[a].n = 0;
        for (int k = 0; k < desc->nv; ++k) {
            const int v = desc->vel[k][a];
            if (v == 0) continue;
            s->sel[a].k[s->sel[a].n] = (unsigned char)k;
            s->sel[a].side[s->sel[a].n] = (v > 0) ? 1 : 2;
            s->sel[a].n++;
        }
    }
    cudaError_t e = cudaStreamCreate(&s->stream);
    {
        // the communication stream outranks the compute stream: the halo exchange that overlaps the inner
        // cells must get SM slots as they free up instead of queueing behind the fused kernel's grid
        int least = 0, greatest = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamDefault, greatest);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_stop);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error(-(int)e, "lbm_sim_create", cudaGetErrorString(e));
        delete s;
        return nullptr;
    }
    return s;
}

// a neighbour rank missed a halo rendezvous (k_wait timed out): every result since then is void
static int check_peers(lbm_sim* s) {
    if (s->wait_err && *(volatile unsigned int*)s->wait_err) {
        char msg[160];
        const unsigned int e = *(volatile unsigned int*)s->wait_err;
        snprintf(msg, sizeof(msg), "rank %d waited more than %.0f s for its %s%s%s neighbour (dead or hung rank?)",
                 s->rank, (double)s->wait_timeout_ns * 1e-9, (e & 1u) ? "left" : "", (e == 3u) ? " and " : "",
                 (e & 2u) ? "right" : "");
        return set_error(-2003, "peer halo timeout", msg);
    }
    return 0;
}

static void free_bc(BcMethod& b) {
    cudaFree(b.istore); cudaFree(b.iload0); cudaFree(b.iload1); cudaFree(b.rhs); cudaFree(b.dist);
    cudaFree(b.istore_odd); cudaFree(b.iload0_odd); cudaFree(b.iload1_odd);
}

static void free_tasks(lbm_sim* s) {
    if (!s->tasks_fn) return;
    cudaFree((void*)s->tasks.block_ptr); cudaFree((void*)s->tasks.code); cudaFree((void*)s->tasks.l0);
    cudaFree((void*)s->tasks.l1); cudaFree((void*)s->tasks.rhs); cudaFree((void*)s->tasks.dist);
    memset(&s->tasks, 0, sizeof(s->tasks));
    s->tasks_fn = nullptr;
}

extern "C" void lbm_sim_destroy(lbm_sim* s) {
    if (!s) return;
    cudaStreamSynchronize(s->stream);
    cudaStreamSynchronize(s->comm_stream);
    if (s->graph) cudaGraphExecDestroy(s->graph);
    if (s->peers_ready) {
        const int nsides = (s->nranks == 2) ? 1 : 2;
        for (int side = 0; side < nsides; ++side) {
            for (int i = 0; i < 2; ++i) if (s->peer_buf[side][i]) cudaIpcCloseMemHandle(s->peer_buf[side][i]);
            if (s->peer_flags[side]) cudaIpcCloseMemHandle(s->peer_flags[side]);
        }
    }
    cudaFree(s->flags);
    if (s->wait_err) cudaFreeHost(s->wait_err);
    // (communicators are shared between the time-step objects of a process and live as long as it does)
    free_tasks(s);
    for (int i = 0; i < 4; ++i) cudaFree(s->xbuf[i]);
    for (int i = 0; i < 2; ++i) {
        if (s->up_buf[i]) cudaFreeHost(s->up_buf[i]);
        if (s->up_event[i]) cudaEventDestroy(s->up_event[i]);
    }
    for (auto& b : s->bcs) free_bc(b);
    cudaFree(s->scratch);
    for (auto e : s->prof_events) cudaEventDestroy(e);
    cudaEventDestroy(s->ev_start); cudaEventDestroy(s->ev_stop);
    cudaEventDestroy(s->ev_ready); cudaEventDestroy(s->ev_comm);
    cudaStreamDestroy(s->stream); cudaStreamDestroy(s->comm_stream);
    delete s;
}

template <typename T>
static cudaError_t upload(T** dst, const T* src, long long n) {
    *dst = nullptr;
    if (!src || n <= 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, (size_t)n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice);
}

static void drop_graph(lbm_sim* s) {
    if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
}

extern "C" int lbm_sim_add_bc(lbm_sim* s, int kind, int64_t ncond, const int64_t* istore, const int64_t* iload0,
                              const int64_t* iload1, const double* rhs, const double* dist, int nlevels,
                              const int64_t* level_ptr, const int* two_phase) {
    if (!s) return ARG_ERROR("null sim");
    if (kind < 0 || kind > LBM_BC_NEUMANN) return ARG_ERROR("unknown boundary kind");
    if (ncond < 0 || (ncond > 0 && (!istore || !iload0))) return ARG_ERROR("null boundary lists");
    const bool two_loads = (kind == LBM_BC_BOUZIDI_BOUNCE_BACK || kind == LBM_BC_BOUZIDI_ANTI_BOUNCE_BACK);
    if (ncond > 0 && two_loads && (!iload1 || !dist)) return ARG_ERROR("Bouzidi needs iload1 and dist");
    if (ncond > 0 && kind != LBM_BC_NEUMANN && !rhs) return ARG_ERROR("rhs missing");
    BcMethod b;
    b.kind = kind;
    b.ncond = ncond;
    {
        cudaError_t e = upload(&b.istore, (const long long*)istore, ncond);
        if (e == cudaSuccess) e = upload(&b.iload0, (const long long*)iload0, ncond);
        if (e == cudaSuccess) e = upload(&b.iload1, (const long long*)iload1, two_loads ? ncond : 0);
        if (e == cudaSuccess) e = upload(&b.rhs, rhs, kind != LBM_BC_NEUMANN ? ncond : 0);
        if (e == cudaSuccess) e = upload(&b.dist, dist, two_loads ? ncond : 0);
        if (e != cudaSuccess) {
            free_bc(b);
            return set_error(-(int)e, "lbm_sim_add_bc", cudaGetErrorString(e));
        }
    }
    if (nlevels <= 0 || !level_ptr) {
        b.level_ptr = {0, (long long)ncond};
        b.two_phase = {kind == LBM_BC_BOUZIDI_BOUNCE_BACK ? 1 : 0};
    } else {
        b.level_ptr.assign(level_ptr, level_ptr + nlevels + 1);
        for (int i = 0; i < nlevels; ++i) b.two_phase.push_back(two_phase ? two_phase[i] : 1);
        if (b.level_ptr.front() != 0 || b.level_ptr.back() != ncond) {
            free_bc(b);
            return ARG_ERROR("level_ptr must span [0, ncond]");
        }
    }
    long long maxlevel = 0;
    for (size_t i = 0; i + 1 < b.level_ptr.size(); ++i) {
        long long n = b.level_ptr[i + 1] - b.level_ptr[i];
        if (n < 0) {
            free_bc(b);
            return ARG_ERROR("level_ptr must be non-decreasing");
        }
        if (n > maxlevel) maxlevel = n;
    }
    if (maxlevel > s->scratch_n) {
        cudaFree(s->scratch);
        s->scratch = nullptr;
        s->scratch_n = 0;
        cudaError_t e = cudaMalloc(&s->scratch, (size_t)maxlevel * sizeof(double));
        if (e != cudaSuccess) {
            s->scratch = nullptr;
            free_bc(b);
            return set_error(-(int)e, "lbm_sim_add_bc: scratch", cudaGetErrorString(e));
        }
        s->scratch_n = maxlevel;
    }
    s->bcs.push_back(b);
    s->bc_groups.clear();
    free_tasks(s);               // a task table describes ALL registered entries
    drop_graph(s);
    return (int)s->bcs.size() - 1;
}

extern "C" int lbm_sim_bc_stale_only(lbm_sim* s, int ibc, int flag) {
    if (!s || ibc < 0 || ibc >= (int)s->bcs.size()) return ARG_ERROR("lbm_sim_bc_stale_only: no such method");
    s->bcs[ibc].stale_only = flag ? 1 : 0;
    drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_set_walls(lbm_sim* s, lbmk_launch_walls_fn launcher, const lbmk_walls* walls) {
    if (!s) return ARG_ERROR("null sim");
    if (walls && !launcher) return ARG_ERROR("lbm_sim_set_walls: launcher missing");
    if (walls && s->tasks_fn) return ARG_ERROR("lbm_sim_set_walls: not combined with lbm_sim_set_tasks");
    if (walls && s->aa_fn) return ARG_ERROR("lbm_sim_set_walls: not combined with in-place streaming");
    if (walls && !(s->wrap_mask & (1 << 2)))
        // the z periodic copy would run on fresh ghosts and overwrite the wall values the kernel stored
        return ARG_ERROR("lbm_sim_set_walls: the fused kernel does not maintain the ghosts of the fastest axis "
                         "(PYLBM_B200_NO_WRAP / NO_ZWRAP, or the axis is too short)");
    if (walls) {
        s->walls = *walls;
        s->walls_fn = launcher;
    } else {
        s->walls_fn = nullptr;
    }
    // ghost values of the current array were produced under the previous setting
    s->ghost_fresh = 0;
    drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_set_tasks(lbm_sim* s, lbmk_launch_tasks_fn launcher, int64_t ntasks, int64_t nblocks,
                                 const int32_t* block_ptr, const uint32_t* code, const int64_t* l0,
                                 const int64_t* l1, const double* dist, const int32_t* ibc, const int64_t* entry,
                                 int ngroups_y, int ngroups_x, int tx) {
    if (!s) return ARG_ERROR("null sim");
    cudaStreamSynchronize(s->stream);
    free_tasks(s);
    drop_graph(s);
    if (ntasks < 0) return 0;
    if (!launcher || nblocks <= 0 || !block_ptr || (ntasks > 0 && (!code || !l0 || !l1 || !dist || !ibc || !entry)))
        return ARG_ERROR("lbm_sim_set_tasks: null argument");
    if (s->walls_fn) return ARG_ERROR("lbm_sim_set_tasks: not combined with lbm_sim_set_walls");
    if (block_ptr[0] != 0 || block_ptr[nblocks] != ntasks) return ARG_ERROR("lbm_sim_set_tasks: block_ptr must span [0, ntasks]");
    std::vector<const double*> rhs((size_t)ntasks, nullptr);
    for (int64_t t = 0; t < ntasks; ++t) {
        if (ibc[t] < 0 || ibc[t] >= (int)s->bcs.size()) return ARG_ERROR("lbm_sim_set_tasks: no such method");
        const BcMethod& b = s->bcs[ibc[t]];
        if (entry[t] < 0 || entry[t] >= b.ncond) return ARG_ERROR("lbm_sim_set_tasks: entry out of range");
        const unsigned kind = code[t] >> 16;
        if ((int)kind != b.kind) return ARG_ERROR("lbm_sim_set_tasks: kind differs from the method's");
        rhs[(size_t)t] = (b.kind == LBM_BC_NEUMANN) ? nullptr : b.rhs + entry[t];
    }
    lbmk_tasks tk;
    memset(&tk, 0, sizeof(tk));
    cudaError_t e = upload((int**)&tk.block_ptr, (const int*)block_ptr, nblocks + 1);
    if (e == cudaSuccess) e = upload((unsigned**)&tk.code, (const unsigned*)code, ntasks);
    if (e == cudaSuccess) e = upload((long long**)&tk.l0, (const long long*)l0, ntasks);
    if (e == cudaSuccess) e = upload((long long**)&tk.l1, (const long long*)l1, ntasks);
    if (e == cudaSuccess) e = upload((const double***)&tk.rhs, (const double* const*)rhs.data(), ntasks);
    if (e == cudaSuccess) e = upload((double**)&tk.dist, dist, ntasks);
    tk.ngroups_y = ngroups_y; tk.ngroups_x = ngroups_x; tk.tx = tx;
    s->tasks = tk;
    s->tasks_fn = launcher;       // (set before a failure is reported so that free_tasks releases the parts)
    if (e != cudaSuccess) {
        free_tasks(s);
        return set_error(-(int)e, "lbm_sim_set_tasks", cudaGetErrorString(e));
    }
    return 0;
}

extern "C" int lbm_sim_set_bc_odd(lbm_sim* s, int ibc, const int64_t* istore, const int64_t* iload0,
                                  const int64_t* iload1) {
    if (!s || ibc < 0 || ibc >= (int)s->bcs.size()) return ARG_ERROR("lbm_sim_set_bc_odd: no such method");
    BcMethod& b = s->bcs[ibc];
    if (b.ncond > 0 && (!istore || !iload0 || (b.iload1 && !iload1))) return ARG_ERROR("lbm_sim_set_bc_odd: null lists");
    cudaFree(b.istore_odd); cudaFree(b.iload0_odd); cudaFree(b.iload1_odd);
    b.istore_odd = b.iload0_odd = b.iload1_odd = nullptr;
    cudaError_t e = upload(&b.istore_odd, (const long long*)istore, b.ncond);
    if (e == cudaSuccess) e = upload(&b.iload0_odd, (const long long*)iload0, b.ncond);
    if (e == cudaSuccess) e = upload(&b.iload1_odd, (const long long*)iload1, b.iload1 ? b.ncond : 0);
    if (e != cudaSuccess) return set_error(-(int)e, "lbm_sim_set_bc_odd", cudaGetErrorString(e));
    drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_set_aa(lbm_sim* s, lbmk_launch_aa_fn launcher) {
    if (!s) return ARG_ERROR("null sim");
    cudaStreamSynchronize(s->stream);
    drop_graph(s);
    if (!launcher) {
        if (s->aa_swapped) return ARG_ERROR("lbm_sim_set_aa: the array is in the swapped layout (odd number of steps)");
        s->aa_fn = nullptr;
        return 0;
    }
    if (s->f != s->fnew) return ARG_ERROR("lbm_sim_set_aa: in-place streaming runs on ONE array (desc.f == desc.fnew)");
    if (s->nranks > 1 && s->peers_ready)
        return ARG_ERROR("lbm_sim_set_aa: on several GPUs in-place streaming uses the NCCL halo, not the peer halo");
    if (s->walls_fn || s->tasks_fn) return ARG_ERROR("lbm_sim_set_aa: not combined with lbm_sim_set_walls / lbm_sim_set_tasks");
    if ((s->wrap_mask | (s->nranks > 1 ? (1 << s->slab_axis) : 0)) != s->d.periodic_mask)
        return ARG_ERROR("lbm_sim_set_aa: the fused kernel must maintain the images of every ghost axis "
                         "(an axis is shorter than four ghost widths, or PYLBM_B200_NO_WRAP is set)");
    s->aa_fn = launcher;
    s->aa_swapped = 0;
    s->ghost_fresh = 0;
    return 0;
}

extern "C" int lbm_sim_set_aa_walls(lbm_sim* s, lbmk_launch_aa_walls_fn launcher, const lbmk_walls* walls) {
    if (!s) return ARG_ERROR("null sim");
    if (!s->aa_fn) return ARG_ERROR("lbm_sim_set_aa_walls: call lbm_sim_set_aa first");
    if (s->aa_swapped) return ARG_ERROR("lbm_sim_set_aa_walls: the array is in the swapped layout");
    if (walls && !launcher) return ARG_ERROR("lbm_sim_set_aa_walls: launcher missing");
    if (walls && !(s->wrap_mask & (1 << 2))) return ARG_ERROR("lbm_sim_set_aa_walls: the fastest axis is not maintained by the kernel");
    cudaStreamSynchronize(s->stream);
    drop_graph(s);
    if (walls) {
        s->walls = *walls;
        s->aa_walls_fn = launcher;
    } else {
        s->aa_walls_fn = nullptr;
    }
    s->aa_even_walled = 0;
    s->ghost_fresh = 0;
    return 0;
}

extern "C" int lbm_sim_aa_phase(lbm_sim* s) { return !s || !s->aa_fn ? -1 : s->aa_swapped; }

extern "C" int lbm_sim_bc_groups(lbm_sim* s, int ngroups, const int* group_ptr) {
    if (!s) return ARG_ERROR("null sim");
    std::vector<int> groups;
    if (ngroups > 0) {
        if (!group_ptr) return ARG_ERROR("lbm_sim_bc_groups: null group_ptr");
        groups.assign(group_ptr, group_ptr + ngroups + 1);
        if (groups.front() != 0 || groups.back() != (int)s->bcs.size())
            return ARG_ERROR("lbm_sim_bc_groups: group_ptr must span all registered methods");
        for (int g = 0; g < ngroups; ++g) {
            const int lo = groups[g], hi = groups[g + 1];
            if (hi <= lo) return ARG_ERROR("lbm_sim_bc_groups: empty group");
            if (hi - lo == 1) continue;
            if (hi - lo > LBM_BC_MAX_SEGMENTS) return ARG_ERROR("lbm_sim_bc_groups: too many methods in a group");
            for (int i = lo; i < hi; ++i) {
                const BcMethod& b = s->bcs[i];
                if (b.ncond > 0 && (b.level_ptr.size() != 2 || b.two_phase[0]))
                    return ARG_ERROR("lbm_sim_bc_groups: a merged method must be one single-phase level");
            }
        }
    }
    s->bc_groups = groups;
    drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_set_rhs(lbm_sim* s, int ibc, const double* rhs_host) {
    if (!s || ibc < 0 || ibc >= (int)s->bcs.size() || !rhs_host) return ARG_ERROR("lbm_sim_set_rhs");
    BcMethod& b = s->bcs[ibc];
    if (b.ncond == 0 || !b.rhs) return 0;
    CUDA_TRY(cudaMemcpyAsync(b.rhs, rhs_host, (size_t)b.ncond * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // rhs_host may be pageable / reused by the caller
    return 0;
}

// right-hand sides recomputed on the device from wall equilibria: rhs[dst[j]] = feq[a[j]] + sign*feq[b[j]]
// (reference: boundary.py:421-427 `rhs = feq[k] - feq[ksym]`, 637-643 with +; one IEEE operation each)
__global__ void k_rhs_from_feq(double* __restrict__ rhs, const double* __restrict__ feq, const long long* __restrict__ a,
                               const long long* __restrict__ b, const long long* __restrict__ dst, double sign,
                               long long count) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    rhs[dst[j]] = __dadd_rn(feq[a[j]], __dmul_rn(sign, feq[b[j]]));
}

extern "C" int lbm_sim_rhs_update(lbm_sim* s, int ibc, int64_t count, const int64_t* a_dev, const int64_t* b_dev,
                                  const int64_t* dst_dev, double sign, const double* feq_dev) {
    if (!s || ibc < 0 || ibc >= (int)s->bcs.size()) return ARG_ERROR("lbm_sim_rhs_update: no such method");
    if (count <= 0) return 0;
    BcMethod& bm = s->bcs[ibc];
    if (!bm.rhs || !a_dev || !b_dev || !dst_dev || !feq_dev || count > bm.ncond)
        return ARG_ERROR("lbm_sim_rhs_update: null argument or more entries than the method has");
    if (sign != 1.0 && sign != -1.0) return ARG_ERROR("lbm_sim_rhs_update: sign must be +1 or -1");
    k_rhs_from_feq<<<(unsigned)((count + 127) / 128), 128, 0, s->stream>>>(
        bm.rhs, feq_dev, (const long long*)a_dev, (const long long*)b_dev, (const long long*)dst_dev, sign, count);
    s->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(-(int)e, "lbm_sim_rhs_update", cudaGetErrorString(e));
    return 0;
}

// Small host block -> device rows on the simulation's stream without stalling the host: the rows go
// through one of two page-locked staging buffers (the host waits only for the copy that used the same
// buffer two uploads ago), so the caller may overwrite `src` as soon as the call returns.
extern "C" int lbm_sim_upload_rows(lbm_sim* s, void* dst_dev, uint64_t dpitch, const void* src_host, uint64_t spitch,
                                   uint64_t width, uint64_t height) {
    if (!s || !dst_dev || !src_host) return ARG_ERROR("lbm_sim_upload_rows: null argument");
    if (width == 0 || height == 0) return 0;
    if (width > spitch || width > dpitch) return ARG_ERROR("lbm_sim_upload_rows: width exceeds a pitch");
    const uint64_t bytes = width * height;
    const int b = s->up_next;
    s->up_next ^= 1;
    if (s->up_event[b]) CUDA_TRY(cudaEventSynchronize(s->up_event[b]));
    else CUDA_TRY(cudaEventCreateWithFlags(&s->up_event[b], cudaEventDisableTiming));
    if (bytes > s->up_bytes[b]) {
        if (s->up_buf[b]) cudaFreeHost(s->up_buf[b]);
        s->up_buf[b] = nullptr;
        s->up_bytes[b] = 0;
        CUDA_TRY(cudaMallocHost(&s->up_buf[b], bytes));
        s->up_bytes[b] = bytes;
    }
    for (uint64_t r = 0; r < height; ++r)
        memcpy((char*)s->up_buf[b] + r * width, (const char*)src_host + r * spitch, width);
    CUDA_TRY(cudaMemcpy2DAsync(dst_dev, dpitch, s->up_buf[b], width, width, height, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaEventRecord(s->up_event[b], s->stream));
    return 0;
}

extern "C" int lbm_sim_set_scalars(lbm_sim* s, const double* scalars, int n) {
    if (!s || n < 0 || n > 32 || (n > 0 && !scalars)) return ARG_ERROR("lbm_sim_set_scalars");
    bool changed = (n != s->d.nscalars);
    for (int i = 0; i < n; ++i) {
        if (s->d.scalars[i] != scalars[i]) changed = true;
        s->d.scalars[i] = scalars[i];
    }
    s->d.nscalars = n;
    if (changed) drop_graph(s);   // kernel arguments are baked into the captured graph
    return 0;
}

// ---- pieces of one step -----------------------------------------------------
// planes of the selected populations <-> one contiguous buffer [population][plane cells]
template <typename S>
__global__ void k_planes(S* __restrict__ f, S* __restrict__ buf, PopSel sel, int side_bit, long long pstride,
                         long long first, long long count, int to_buf) {
    // `first` = element position of the first plane inside population 0 (lead + plane * stride)
    int nsel = 0;
    unsigned char ks[64];
    for (int i = 0; i < sel.n; ++i)
        if (sel.side[i] & side_bit) ks[nsel++] = sel.k[i];
    const long long total = (long long)nsel * count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / count, j = i - p * count;
        S* a = f + (long long)ks[p] * pstride + first + j;
        if (to_buf) buf[i] = *a; else *a = buf[i];
    }
}

static cudaError_t launch_planes(lbm_sim* s, void* f, void* buf, int side_bit, long long first, long long count,
                                 int nsel, int to_buf, cudaStream_t st) {
    const long long total = (long long)nsel * count;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    const int a = s->slab_axis;
    if (s->d.storage == LBM_STORAGE_F64)
        k_planes<double><<<(unsigned)blocks, 256, 0, st>>>((double*)f, (double*)buf, s->sel[a], side_bit,
                                                            s->d.grid.pstride, first, count, to_buf);
    else
        k_planes<float><<<(unsigned)blocks, 256, 0, st>>>((float*)f, (float*)buf, s->sel[a], side_bit,
                                                           s->d.grid.pstride, first, count, to_buf);
    s->launches += 1;
    return cudaGetLastError();
}

static int exchange_slabs(lbm_sim* s, void* f, cudaStream_t st, bool reverse = false) {
    // planes [w, 2w) go to the left neighbour's right ghost, planes [n-2w, n-w) to the right
    // neighbour's left ghost (periodic ring, like the reference's Cartesian communicator).  Only the
    // populations that enter the neighbour through that face travel (sign-matched, PopSel), packed into
    // ONE message per direction: 2 sends + 2 receives per step instead of one pair per population
    // (20 messages of 2.2 MB for D3Q19 at 512^3 cost 0.54 ms of per-message overhead on 8 GPUs).
    const lbmk_grid& g = s->d.grid;
    const int a = s->slab_axis;
    const int w = s->d.vmax[a];
    if (w <= 0) return 0;
    const long long n = g.n[a];
    const long long stride = (a == 0) ? (long long)g.n[1] * g.pitch : (a == 1) ? g.pitch : 1;
    const long long count = (long long)w * stride;
    const int left = (s->rank + s->nranks - 1) % s->nranks, right = (s->rank + 1) % s->nranks;
    const size_t esz = (s->d.storage == LBM_STORAGE_F64) ? 8 : 4;
    const int dtype = (s->d.storage == LBM_STORAGE_F64) ? NCCL_FLOAT64 : NCCL_FLOAT32;
    const PopSel& sel = s->sel[a];
    int n_neg = 0, n_pos = 0;          // populations moving in -axis (side bit 2) / +axis (side bit 1)
    for (int i = 0; i < sel.n; ++i) {
        if (sel.side[i] & 2) ++n_neg;
        if (sel.side[i] & 1) ++n_pos;
    }
    // REVERSE exchange (in-place streaming, after an even step): the ghost planes, where the even kernel
    // deposited the populations that left through the slab faces, travel to the neighbours' interior
    // planes -- the low ghost layer holds the slots of the populations moving in +axis, and it goes to the
    // LEFT neighbour's planes [n-2w, n-w); the high one (-axis) to the right neighbour's planes [w, 2w).
    const int bit_left = reverse ? 1 : 2, bit_right = reverse ? 2 : 1;        // populations sent left / right
    const int n_left = reverse ? n_pos : n_neg, n_right = reverse ? n_neg : n_pos;
    const long long src_left = reverse ? 0 : (long long)w, src_right = reverse ? n - w : n - 2 * w;
    const long long dst_from_right = reverse ? n - 2 * w : n - w, dst_from_left = reverse ? (long long)w : 0;
    // buffers: [0] send to the left, [1] send to the right, [2] recv from the right, [3] recv from the left
    const size_t need[4] = {(size_t)n_left * count * esz, (size_t)n_right * count * esz,
                            (size_t)n_left * count * esz, (size_t)n_right * count * esz};
    for (int i = 0; i < 4; ++i) {
        if (need[i] > s->xbuf_bytes[i]) {
            if (s->xbuf[i]) cudaFree(s->xbuf[i]);
            s->xbuf[i] = nullptr;
            s->xbuf_bytes[i] = 0;
            CUDA_TRY(cudaMalloc(&s->xbuf[i], need[i]));
            s->xbuf_bytes[i] = need[i];
        }
    }
    cudaError_t e = launch_planes(s, f, s->xbuf[0], bit_left, g.lead + src_left * stride, count, n_left, 1, st);
    if (e == cudaSuccess) e = launch_planes(s, f, s->xbuf[1], bit_right, g.lead + src_right * stride, count, n_right, 1, st);
    if (e != cudaSuccess) return set_error(-(int)e, "halo pack", cudaGetErrorString(e));
    NCCL_TRY(g_nccl.GroupStart());
    // receives first from the right, then from the left: with 2 ranks both neighbours are the same peer
    // and messages are matched in posting order (what the right neighbour sends to ITS left is what it
    // packed with bit_left)
    if (n_left) NCCL_TRY(g_nccl.Recv(s->xbuf[2], (size_t)n_left * count, dtype, right, s->comm, st));
    if (n_right) NCCL_TRY(g_nccl.Recv(s->xbuf[3], (size_t)n_right * count, dtype, left, s->comm, st));
    if (n_left) NCCL_TRY(g_nccl.Send(s->xbuf[0], (size_t)n_left * count, dtype, left, s->comm, st));
    if (n_right) NCCL_TRY(g_nccl.Send(s->xbuf[1], (size_t)n_right * count, dtype, right, s->comm, st));
    NCCL_TRY(g_nccl.GroupEnd());
    s->launches += 1;
    // forward: the high ghost layer is only read for populations moving in -axis, the low one for +axis
    e = launch_planes(s, f, s->xbuf[2], bit_left, g.lead + dst_from_right * stride, count, n_left, 0, st);
    if (e == cudaSuccess) e = launch_planes(s, f, s->xbuf[3], bit_right, g.lead + dst_from_left * stride, count, n_right, 0, st);
    if (e != cudaSuccess) return set_error(-(int)e, "halo unpack", cudaGetErrorString(e));
    return 0;
}

static int apply_bc_method(lbm_sim* s, BcMethod& b, void* f, cudaStream_t st) {
    // the previous fused launch stored these values (in place: the even step, into the cells' own slots)
    if (b.stale_only && (s->ghost_fresh || (s->aa_fn && s->aa_swapped && s->aa_even_walled))) return 0;
    const bool odd = s->aa_fn && s->aa_swapped;
    if (odd && b.ncond > 0 && !b.istore_odd) return ARG_ERROR("in-place streaming: odd-step lists missing (lbm_sim_set_bc_odd)");
    const long long* is_ = odd ? b.istore_odd : b.istore;
    const long long* l0_ = odd ? b.iload0_odd : b.iload0;
    const long long* l1_ = odd ? b.iload1_odd : b.iload1;
    for (size_t l = 0; l + 1 < b.level_ptr.size(); ++l) {
        const long long o = b.level_ptr[l], n = b.level_ptr[l + 1] - o;
        if (n <= 0) continue;
        cudaError_t e =
            (s->d.storage == LBM_STORAGE_F64)
                ? launch_bc<double>(b.kind, (double*)f, n, is_ + o, l0_ + o,
                                    l1_ ? l1_ + o : nullptr, b.rhs ? b.rhs + o : nullptr,
                                    b.dist ? b.dist + o : nullptr, s->scratch, b.two_phase[l], st, &s->launches)
                : launch_bc<float>(b.kind, (float*)f, n, is_ + o, l0_ + o,
                                   l1_ ? l1_ + o : nullptr, b.rhs ? b.rhs + o : nullptr,
                                   b.dist ? b.dist + o : nullptr, s->scratch, b.two_phase[l], st, &s->launches);
        if (e != cudaSuccess) return set_error(-(int)e, "boundary kernel", cudaGetErrorString(e));
    }
    return 0;
}

static int apply_bcs(lbm_sim* s, void* f, cudaStream_t st) {
    if (s->bc_groups.empty()) {
        for (auto& b : s->bcs) {
            int rc = apply_bc_method(s, b, f, st);
            if (rc) return rc;
        }
        return 0;
    }
    for (size_t g = 0; g + 1 < s->bc_groups.size(); ++g) {
        const int lo = s->bc_groups[g], hi = s->bc_groups[g + 1];
        if (hi - lo == 1) {
            int rc = apply_bc_method(s, s->bcs[lo], f, st);
            if (rc) return rc;
            continue;
        }
        BcSegments segs;
        segs.n = 0;
        segs.total = 0;
        for (int i = lo; i < hi; ++i) {
            BcMethod& b = s->bcs[i];
            if (b.ncond <= 0 || (b.stale_only && (s->ghost_fresh || (s->aa_fn && s->aa_swapped && s->aa_even_walled))))
                continue;
            BcSegment& sg = segs.seg[segs.n++];
            sg.begin = segs.total;
            sg.kind = b.kind;
            const bool odd = s->aa_fn && s->aa_swapped;
            if (odd && !b.istore_odd) return ARG_ERROR("in-place streaming: odd-step lists missing (lbm_sim_set_bc_odd)");
            sg.istore = odd ? b.istore_odd : b.istore;
            sg.iload0 = odd ? b.iload0_odd : b.iload0;
            sg.iload1 = odd ? b.iload1_odd : b.iload1;
            sg.rhs = b.rhs; sg.dist = b.dist;
            segs.total += b.ncond;
        }
        if (segs.total == 0) continue;
        const unsigned blocks = (unsigned)((segs.total + 127) / 128);
        cudaError_t e = (s->d.storage == LBM_STORAGE_F64)
                            ? launch_k(k_bc_multi<double>, dim3(blocks), dim3(128), st, g_pdl, (double*)f, segs)
                            : launch_k(k_bc_multi<float>, dim3(blocks), dim3(128), st, g_pdl, (float*)f, segs);
        s->launches += 1;
        if (e != cudaSuccess) return set_error(-(int)e, "boundary kernel", cudaGetErrorString(e));
    }
    return 0;
}

static int ghost_update(lbm_sim* s, void* f, cudaStream_t st) {
    const lbmk_grid& g = s->d.grid;
    int mask = s->d.periodic_mask;
    if (s->nranks > 1) {
        if (s->peers_ready && s->ghost_fresh) {
            // the neighbours stored their slab-face images into my ghost planes during their previous
            // fused kernel: just wait for both of them to have finished it (once per array state:
            // lbm_sim_boundary_condition may already have done it)
            if (!s->waited) {
                launch_k(k_wait, dim3(1), dim3(32), st, g_pdl, s->flags, s->wait_err_dev, s->wait_timeout_ns);
                s->launches += 1;
                s->waits += 1;
                s->waited = 1;
            }
        } else if (s->xchg_inflight && s->ghost_fresh) {
            // NCCL halo, steady state: the slab-face planes of this array were exchanged on the
            // communication stream while the previous step computed its inner cells
            CUDA_TRY(cudaStreamWaitEvent(st, s->ev_comm, 0));
            s->xchg_inflight = 0;
        } else {
            if (s->xchg_inflight) {               // stale ghosts (outside write): redo it after the one in flight
                CUDA_TRY(cudaStreamWaitEvent(st, s->ev_comm, 0));
                s->xchg_inflight = 0;
            }
            int rc = exchange_slabs(s, f, st);
            if (rc) return rc;
            if (s->peers_ready && s->waits < s->signals) {
                // a signal of the previous step is still pending (the ghosts were invalidated from
                // outside): consume it so that the arrival counters stay in step
                launch_k(k_wait, dim3(1), dim3(32), st, g_pdl, s->flags, s->wait_err_dev, s->wait_timeout_ns);
                s->launches += 1;
                s->waits += 1;
            }
        }
        mask &= ~(1 << s->slab_axis);
    }
    if (s->ghost_fresh) mask &= ~s->wrap_mask;   // images were written by the previous fused launch
    cudaError_t e = periodic_axes(f, g, s->sel, s->d.storage, s->d.vmax, mask, 0, g.n[0], st, &s->launches);
    if (e != cudaSuccess) return set_error(-(int)e, "periodic update", cudaGetErrorString(e));
    return 0;
}

static int one_step(lbm_sim* s, void* f, void* fnew, double t, cudaStream_t st) {
    if (s->aa_fn) {
        // in-place streaming (AA pattern).  Even step: ghost update, boundary methods, gather + scatter
        // back (the array is swapped afterwards).  Odd step: boundary methods on the transformed lists
        // -- no ghost update, the even step stored the wrapped images itself -- then the local kernel,
        // which leaves the natural layout and the periodic images of the next even step.
        int rc = 0;
        if (!s->aa_swapped) {
            rc = ghost_update(s, f, st);
            if (rc) return rc;
        }
        rc = apply_bcs(s, f, st);
        if (rc) return rc;
        double scal[32];
        for (int i = 0; i < s->d.nscalars; ++i) scal[i] = s->d.scalars[i];
        if (s->d.t_index >= 0 && s->d.t_index < s->d.nscalars) scal[s->d.t_index] = t;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (s->profile) {
            if (s->prof_used + 2 > s->prof_events.size()) {
                cudaEvent_t a, b;
                if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess)
                    return set_error(-2002, "profile", "cannot create events");
                s->prof_events.push_back(a);
                s->prof_events.push_back(b);
            }
            ev0 = s->prof_events[s->prof_used];
            ev1 = s->prof_events[s->prof_used + 1];
            s->prof_used += 2;
            cudaEventRecord(ev0, st);
        }
        lbmk_grid g = s->d.grid;
        g.wrap = s->wrap_mask | (g_pdl ? LBMK_WRAP_PDL : 0);
        if (s->aa_walls_fn)
            rc = s->aa_walls_fn(f, &g, scal, s->aa_swapped, &s->walls, (void*)st);
        else
            rc = s->aa_fn(f, &g, scal, s->aa_swapped, (void*)st);
        if (rc) return set_error(rc, "in-place kernel launch", cudaGetErrorString((cudaError_t)(-rc)));
        s->aa_even_walled = (!s->aa_swapped && s->aa_walls_fn) ? 1 : 0;
        if (ev1) cudaEventRecord(ev1, st);
        s->launches += 1;
        if (!s->aa_swapped && s->nranks > 1) {
            // x-slabs: what left through the slab faces sits in my ghost planes; the neighbours read it in
            // their own swapped slots during the odd step
            rc = exchange_slabs(s, f, st, true);
            if (rc) return rc;
        }
        s->aa_swapped ^= 1;
        s->ghost_fresh = s->aa_swapped ? 0 : 1;     // the odd kernel wrote the images of the next even step
        return 0;
    }
    int rc = ghost_update(s, f, st);
    if (rc) return rc;
    if (!s->tasks_fn) {          // with a task table the fused kernel evaluates the entries itself
        rc = apply_bcs(s, f, st);
        if (rc) return rc;
    }
    double scal[32];
    for (int i = 0; i < s->d.nscalars; ++i) scal[i] = s->d.scalars[i];
    if (s->d.t_index >= 0 && s->d.t_index < s->d.nscalars) scal[s->d.t_index] = t;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (s->profile) {
        if (s->prof_used + 2 > s->prof_events.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess)
                return set_error(-2002, "profile", "cannot create events");
            s->prof_events.push_back(a);
            s->prof_events.push_back(b);
        }
        ev0 = s->prof_events[s->prof_used];
        ev1 = s->prof_events[s->prof_used + 1];
        s->prof_used += 2;
        cudaEventRecord(ev0, st);
    }
    lbmk_grid g = s->d.grid;
    g.wrap = s->wrap_mask | (g_pdl ? LBMK_WRAP_PDL : 0);
    lbmk_peers pr;
    if (s->peers_ready) {
        const int which = (fnew == s->buf[0]) ? 0 : 1;   // all ranks swap A/B in lockstep
        pr.lo = s->peer_buf[0][which];
        pr.hi = s->peer_buf[1][which];
        pr.pstride_lo = s->peer_pstride[0];
        pr.pstride_hi = s->peer_pstride[1];
        pr.nin_lo = s->peer_nin_lo;
    }
    auto launch_fused = [&](const lbmk_grid& gg) -> int {
        if (s->tasks_fn)
            return s->tasks_fn(f, fnew, &gg, scal, s->peers_ready ? &pr : nullptr, &s->tasks, (void*)st);
        if (s->walls_fn)
            return s->walls_fn(f, fnew, &gg, scal, s->peers_ready ? &pr : nullptr, &s->walls, (void*)st);
        if (s->peers_ready) return s->d.one_time_step_peers(f, fnew, &gg, scal, &pr, (void*)st);
        return s->d.one_time_step(f, fnew, &gg, scal, (void*)st);
    };
    const int ax = s->slab_axis, wx = s->d.vmax[ax];
    const bool split = s->nranks > 1 && !s->peers_ready && s->overlap && s->comm && wx > 0 &&
                       g.hi[ax] - g.lo[ax] > 4 * wx && !(s->tasks_fn && ax != 0);
    if (split) {
        // NCCL halo with overlap: the cells within `wx` of the slab faces first, then their planes travel
        // on the communication stream (they are the ghost planes of the neighbours' NEXT step) while
        // the inner cells are computed.  BC kernels of the next step run after the exchange, in the
        // reference's order (simulation.py:384-390).
        lbmk_grid lo = g, hi = g, in = g;
        lo.hi[ax] = g.lo[ax] + wx;
        hi.lo[ax] = g.hi[ax] - wx;
        in.lo[ax] = g.lo[ax] + wx;
        in.hi[ax] = g.hi[ax] - wx;
        rc = launch_fused(lo);
        if (rc == 0) rc = launch_fused(hi);
        if (rc == 0) {
            CUDA_TRY(cudaEventRecord(s->ev_ready, st));
            CUDA_TRY(cudaStreamWaitEvent(s->comm_stream, s->ev_ready, 0));
            int rx = exchange_slabs(s, fnew, s->comm_stream);
            if (rx) return rx;
            CUDA_TRY(cudaEventRecord(s->ev_comm, s->comm_stream));
            s->xchg_inflight = 1;
            rc = launch_fused(in);
            s->launches += 2;
        }
    } else {
        rc = launch_fused(g);
    }
    if (rc) return set_error(rc, "one_time_step kernel launch", cudaGetErrorString((cudaError_t)(-rc)));
    if (ev1) cudaEventRecord(ev1, st);
    s->launches += 1;
    if (s->peers_ready) {
        launch_k(k_signal, dim3(1), dim3(32), st, g_pdl, s->peer_flags[0] + 1, s->peer_flags[1] + 0);
        s->launches += 1;
        s->signals += 1;
        s->waited = 0;
    }
    s->ghost_fresh = 1;   // fnew (the next f) now carries its periodic images
    return 0;
}

static int build_graph(lbm_sim* s) {
    drop_graph(s);
    cudaGraph_t graph = nullptr;
    const int64_t before = s->launches;
    const long long sig_before = s->signals, wait_before = s->waits;
    CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    int rc = one_step(s, s->f, s->fnew, s->t, s->stream);
    if (rc == 0) rc = one_step(s, s->fnew, s->f, s->t, s->stream);
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    s->graph_launches = s->launches - before;
    s->graph_signals = s->signals - sig_before;
    s->launches = before;
    s->signals = sig_before;      // nothing was executed during the capture
    s->waits = wait_before;
    s->waited = 0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return set_error(-(int)e, "cudaStreamEndCapture", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_error(-(int)e, "cudaGraphInstantiate", cudaGetErrorString(e));
    s->graph_f = s->f;
    return 0;
}

extern "C" int lbm_sim_step(lbm_sim* s, int nsteps) {
    if (!s || nsteps < 0) return ARG_ERROR("lbm_sim_step");
    if (int rc = check_peers(s)) return rc;
    int done = 0;
    const bool graph_ok = s->use_graph && s->d.t_index < 0 && (s->nranks == 1 || s->peers_ready) && !s->profile;
    while (done < nsteps) {
        // pairs of steps (f -> fnew -> f) are replayed from a CUDA graph once the ghost layers are
        // maintained by the fused kernel itself (no copy kernels / NCCL calls inside the capture)
        if (graph_ok && s->ghost_fresh && !s->waited && nsteps - done >= 2 && !(s->aa_fn && s->aa_swapped)) {
            if (!s->graph || s->graph_f != s->f) {
                int rc = build_graph(s);
                if (rc) return rc;
            }
            CUDA_TRY(cudaGraphLaunch(s->graph, s->stream));
            s->launches += s->graph_launches;
            s->signals += s->graph_signals;
            s->waits += s->graph_signals;
            s->t += 2 * s->d.dt;
            s->nt += 2;
            done += 2;
            continue;
        }
        int rc = one_step(s, s->f, s->fnew, s->t, s->stream);
        if (rc) return rc;
        void* tmp = s->f; s->f = s->fnew; s->fnew = tmp;
        s->t += s->d.dt;
        s->nt += 1;
        done += 1;
    }
    return 0;
}

extern "C" int lbm_sim_boundary_condition(lbm_sim* s) {
    if (!s) return ARG_ERROR("null sim");
    if (!(s->aa_fn && s->aa_swapped)) {
        int rc = ghost_update(s, s->f, s->stream);
        if (rc) return rc;
    }
    return apply_bcs(s, s->f, s->stream);
}

extern "C" int lbm_sim_sync(lbm_sim* s) {
    if (!s) return ARG_ERROR("null sim");
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->comm_stream));
    return check_peers(s);
}

extern "C" int lbm_sim_state(lbm_sim* s, void** f, void** fnew, double* t, int64_t* nt) {
    if (!s) return ARG_ERROR("null sim");
    if (f) *f = s->f;
    if (fnew) *fnew = s->fnew;
    if (t) *t = s->t;
    if (nt) *nt = s->nt;
    return 0;
}

extern "C" int lbm_sim_set_state(lbm_sim* s, void* f, void* fnew, double t) {
    if (!s || !f || !fnew) return ARG_ERROR("lbm_sim_set_state");
    s->f = f; s->fnew = fnew; s->t = t;
    s->ghost_fresh = 0;
    return 0;
}

extern "C" int lbm_sim_invalidate_ghosts(lbm_sim* s) {
    if (!s) return ARG_ERROR("null sim");
    s->ghost_fresh = 0;
    return 0;
}

extern "C" int lbm_sim_use_graph(lbm_sim* s, int enable) {
    if (!s) return ARG_ERROR("null sim");
    s->use_graph = enable ? 1 : 0;
    if (!enable) drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_timer_start(lbm_sim* s) {
    if (!s) return ARG_ERROR("null sim");
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaEventRecord(s->ev_start, s->stream));
    return 0;
}

extern "C" int lbm_sim_timer_stop(lbm_sim* s, float* ms) {
    if (!s || !ms) return ARG_ERROR("lbm_sim_timer_stop");
    CUDA_TRY(cudaEventRecord(s->ev_stop, s->stream));
    CUDA_TRY(cudaEventSynchronize(s->ev_stop));
    CUDA_TRY(cudaEventElapsedTime(ms, s->ev_start, s->ev_stop));
    return check_peers(s);
}

extern "C" int lbm_sim_profile(lbm_sim* s, int enable) {
    if (!s) return ARG_ERROR("null sim");
    s->profile = enable ? 1 : 0;
    s->prof_used = 0;
    return 0;
}

extern "C" int lbm_sim_profile_read(lbm_sim* s, double* fused_ms, int64_t* nlaunch) {
    if (!s || !fused_ms || !nlaunch) return ARG_ERROR("lbm_sim_profile_read");
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    double total = 0.0;
    for (size_t i = 0; i + 1 < s->prof_used; i += 2) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->prof_events[i], s->prof_events[i + 1]));
        total += ms;
    }
    *fused_ms = total;
    *nlaunch = (int64_t)(s->prof_used / 2);
    s->prof_used = 0;
    return 0;
}

extern "C" int64_t lbm_sim_launch_count(lbm_sim* s) { return s ? s->launches : 0; }
extern "C" void* lbm_sim_stream(lbm_sim* s) { return s ? (void*)s->stream : nullptr; }

struct IpcBlob {
    cudaIpcMemHandle_t buf[2];
    cudaIpcMemHandle_t flags;
    long long pstride;
    int nin;
    int pad;
};
static_assert(sizeof(IpcBlob) <= LBM_IPC_BLOB_BYTES, "IPC blob too large");

extern "C" int lbm_sim_ipc_export(lbm_sim* s, void* blob256) {
    if (!s || !blob256) return ARG_ERROR("lbm_sim_ipc_export");
    if (!s->flags) {
        CUDA_TRY(cudaMalloc(&s->flags, 4 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(s->flags, 0, 4 * sizeof(unsigned long long)));
        CUDA_TRY(cudaHostAlloc((void**)&s->wait_err, sizeof(unsigned int), cudaHostAllocMapped));
        *s->wait_err = 0u;
        CUDA_TRY(cudaHostGetDevicePointer((void**)&s->wait_err_dev, s->wait_err, 0));
        const char* env = getenv("PYLBM_B200_HALO_TIMEOUT_S");
        if (env && atof(env) > 0.0) s->wait_timeout_ns = (unsigned long long)(atof(env) * 1e9);
        CUDA_TRY(cudaDeviceSynchronize());
    }
    IpcBlob b;
    memset(&b, 0, sizeof(b));
    CUDA_TRY(cudaIpcGetMemHandle(&b.buf[0], s->buf[0]));
    CUDA_TRY(cudaIpcGetMemHandle(&b.buf[1], s->buf[1]));
    CUDA_TRY(cudaIpcGetMemHandle(&b.flags, s->flags));
    b.pstride = s->d.grid.pstride;
    b.nin = s->d.grid.n[s->slab_axis] - 2 * s->d.vmax[s->slab_axis];
    memset(blob256, 0, LBM_IPC_BLOB_BYTES);
    memcpy(blob256, &b, sizeof(b));
    return 0;
}

extern "C" int lbm_sim_ipc_open(lbm_sim* s, const void* left_blob, const void* right_blob) {
    if (!s || !left_blob || !right_blob) return ARG_ERROR("lbm_sim_ipc_open");
    if (!s->d.one_time_step_peers) return ARG_ERROR("the kernel library has no lbmk_one_time_step_peers");
    if (s->nranks < 2 || !s->flags) return ARG_ERROR("lbm_sim_ipc_open needs lbm_sim_comm_init and lbm_sim_ipc_export first");
    if (s->aa_fn) return ARG_ERROR("lbm_sim_ipc_open: in-place streaming uses the NCCL halo");
    const void* blobs[2] = {left_blob, right_blob};
    const bool same = (s->nranks == 2);   // both neighbours are the same process: open its handles once
    for (int side = 0; side < 2; ++side) {
        IpcBlob b;
        memcpy(&b, blobs[side], sizeof(b));
        if (side == 1 && same) {
            for (int i = 0; i < 2; ++i) s->peer_buf[1][i] = s->peer_buf[0][i];
            s->peer_flags[1] = s->peer_flags[0];
        } else {
            for (int i = 0; i < 2; ++i)
                CUDA_TRY(cudaIpcOpenMemHandle(&s->peer_buf[side][i], b.buf[i], cudaIpcMemLazyEnablePeerAccess));
            CUDA_TRY(cudaIpcOpenMemHandle((void**)&s->peer_flags[side], b.flags, cudaIpcMemLazyEnablePeerAccess));
        }
        s->peer_pstride[side] = b.pstride;
        if (side == 0) s->peer_nin_lo = b.nin;
    }
    s->peers_ready = 1;
    s->ghost_fresh = 0;
    drop_graph(s);
    return 0;
}

extern "C" int lbm_sim_comm_init(lbm_sim* s, int rank, int nranks, const void* id128) {
    if (!s || nranks < 1 || rank < 0 || rank >= nranks) return ARG_ERROR("lbm_sim_comm_init");
    if (s->aa_fn) return ARG_ERROR("lbm_sim_comm_init: call it before lbm_sim_set_aa");
    s->rank = rank;
    s->nranks = nranks;
    if (nranks == 1) return 0;
    if (!id128) return ARG_ERROR("null unique id");
    int rc = nccl_load();
    if (rc) return rc;
    // One communicator per unique id and process, kept for the life of the process: an ncclUniqueId can
    // bootstrap only one communicator, and a process builds several time-step objects over the same
    // ranks one after the other (bench.py: parity cases, the headline workload, the other configurations).
    const std::string key((const char*)id128, 128);
    auto it = g_comms.find(key);
    if (it != g_comms.end()) {
        if (it->second.rank != rank || it->second.nranks != nranks)
            return ARG_ERROR("lbm_sim_comm_init: this unique id already belongs to a communicator of another shape");
        s->comm = it->second.comm;
    } else {
        nccl_uid id;
        memcpy(&id, id128, 128);
        NCCL_TRY(g_nccl.CommInitRank(&s->comm, nranks, id, rank));
        g_comms[key] = CommEntry{s->comm, rank, nranks};
    }
    if (getenv("PYLBM_B200_NO_OVERLAP")) s->overlap = 0;   // debugging aid: exchange on the compute stream
    s->wrap_mask &= ~(1 << s->slab_axis);   // the slab axis is exchanged between ranks, not wrapped
    s->ghost_fresh = 0;
    drop_graph(s);
    return 0;
}
