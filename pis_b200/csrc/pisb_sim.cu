// pisb_sim.cu -- handle, device memory, launch sequences and the C ABI of libpisb200.so.
// Kernels live in pisb_kernels.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <ctime>
#include <string>
#include <vector>

#include "pisb200.h"
#include "pisb_kernels.cuh"
#include "pisb_multi.cuh"
#include "pisb_npt.cuh"
#include "pisb_velocity.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler is attached

using namespace pisb;

namespace {

thread_local std::string g_create_error;

std::string fmt(const char *f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

// max{t : sqrt(t) <= rc}: with IEEE sqrt monotone, `sqrt(r2) > rc` <=> `r2 > t` exactly.
double sqrt_threshold(double rc) {
    if (!(rc > 0.0) || !std::isfinite(rc)) return rc > 0.0 ? rc : 0.0;
    double t = rc * rc;
    while (std::sqrt(t) > rc) t = std::nextafter(t, 0.0);
    while (std::sqrt(std::nextafter(t, INFINITY)) <= rc) t = std::nextafter(t, INFINITY);
    return t;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
};

}  // namespace

struct pisb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // overlaps the position read-back of the host-buffer step with the force kernel
    cudaEvent_t ev_pos = nullptr;
    cudaStream_t up_stream = nullptr;    // chunked host-buffer step: host->device copies run ahead of the drift kernels
    cudaStream_t dl_stream = nullptr;    // multi-GPU: the asynchronous owned-atom download (copy_stream carries the per-step flag there)
    cudaEvent_t ev_dl = nullptr;
    DevBuf<int> dl_ids;
    // the frame of pisb_download_owned_begin travels in pieces, a budget per step (pump_download): a frame in flight delays
    // every host-visible word of the step loop that queues behind it on the link
    struct DlPiece {
        void *dst;
        const void *src;
        size_t bytes;
    };
    std::vector<DlPiece> dl_queue;
    size_t dl_next = 0, dl_budget = 0;
    int dl_spread_steps = 8;  // option: steps a frame is spread over
    std::vector<cudaEvent_t> ev_chunk;   // ... one arrival / one departure event per chunk
    int host_pipeline = 1;               // option "host_pipeline": 0 = whole-array copies (the unpipelined sequence)
    int host_chunk_atoms = 0;            // option "host_chunk_atoms": atoms per chunk, 0 = auto (n/8, at least 65536)
    std::string err;

    // potential
    int n_types = 0;
    std::vector<double> mass, eps, sigma, rcut;
    std::vector<unsigned char> present;
    int shift = 1;
    double skin = 0.0;
    double max_rcut = 0.0;
    std::vector<PairDev> pairs;
    std::vector<PairF> pairsf;
    BoxF boxf{};
    int force_variant = 0;  // 0 = auto (v3 when orthorhombic + fully periodic, else v1), 1 = v1, 2 = v2, 3 = v3
    int build_variant = 0;
    int cell_div = 0;  // cells per list cutoff per dimension: 0 = auto, 1 = reference-sized cells, 2 = half-size cells
    int fuse_vv = 1;   // option "fuse_vv": NVE batches run k_force_vv (force + kick + drift in one launch) when the default force kernel applies

    // box / grid
    BoxDev box{};
    bool have_box = false;
    Grid grid{};
    bool grid_ok = false;

    // atoms
    int n = 0, npad = 0, kcap = 0;
    bool have_atoms = false;
    bool list_valid = false;   // a list exists for the current atom set / box
    bool forces_current = false;
    int kcap_user = 0;
    int n_capacity_growths = 0;  // grow_list_if_close: times the capacity was raised ahead of an overflow

    // device arrays
    DevBuf<double4> xt, s_xt;
    DevBuf<float4> xf, xf2;  // xf2: the other FP32 shadow buffer of the fused force+integrator step (xt alternates with s_xt)
    int xt_flag = FLAG_UNWRAPPED_A, s_xt_flag = FLAG_UNWRAPPED_B;  // flags[] word of each position buffer ("holds unwrapped coordinates")
    DevBuf<float> xp;  // pair-packed FP32 positions (k_build_list_v3), rewritten at every rebuild
    DevBuf<PairF> tablef_d;
    DevBuf<double> v[3], f[3], g[3], xb[3], s_v[3], s_f[3];
    DevBuf<int> id, s_id, slot_of_id, cell_of, order, nnbr, nbr;
    DevBuf<int> cell_count, cell_start, tile_sum;
    DevBuf<double> mass_d, partials, st_pos, st_vel, st_frc;
    DevBuf<double> dl_pos, dl_vel, dl_frc;  // snapshot of an asynchronous download (pisb_download_begin / _end)
    bool dl_pending = false;
    DevBuf<int> st_types;
    DevBuf<PairDev> table_d;
    DevBuf<NhcDev> nhc_d;
    DevBuf<double> nhc_energy_d;
    DevBuf<double> npt_tensors_d;
    DevBuf<double> vel_sums_d;  // pisb_start_velocities: total mass, momentum, atom count, kinetic energy
    double *h_npt = nullptr;         // pinned: 19 reduced tensor sums + thermostat energy
    double *h_energy = nullptr;      // pinned: thermostat energy per step of an NVT batch (a pageable target would make every
    size_t h_energy_cap = 0;         //         piece's read-back a synchronous copy)
    double skin_half2_override = -1.0;  // NPT: skin trigger threshold reduced by the accumulated box strain
    DevBuf<pisb_thermo> thermo_d;
    int *flags = nullptr;
    unsigned int *ticket = nullptr;
    size_t ticket_cap = 0;  // words: 1 + one per group of 64 blocks (block_reduce_finalize)
    int *h_flags = nullptr;          // pinned; [FLAG_COUNT] = sequence number of k_publish_flags
    int pub_seq = 0;
    int *h_pub = nullptr;            // pinned, 16 ints behind h_flags: one published record (k_publish_words) + its sequence word
    bool table_on_device = false;  // table_d holds the pair table (setup_filter)
    bool list_checked = false;     // ensure_list has confirmed the list for the CURRENT positions and the device's REBUILD flag is clear
                                   // (kept by the NVE / NVT batches, which end on a kick; cleared by uploads, host-buffer steps, NPT)
    int64_t missing_type_pairs = 0;  // populated type pairs without a potential (count_missing_pairs)
    pisb_thermo *h_thermo = nullptr; // pinned
    size_t h_thermo_cap = 0;
    int64_t device_bytes = 0;

    // multi-GPU (spatial decomposition; see pisb_multi.cuh)
    bool multi = false;
    Decomp dc{};
    ncclComm_t comm = nullptr;
    int n_own = 0, n_ghost = 0, ncap_atoms = 0;
    DevBuf<int> m_dest, m_pig, m_cnt, m_allcnt, m_off, send_idx, ghost_slot, newslot, flag_in;
    DevBuf<double> mig_send, mig_recv;
    DevBuf<double4> halo_send, halo_recv;
    std::vector<int> gs_cnt, gr_cnt, gs_off, gr_off;  // per-peer ghost send/recv counts and offsets
    int send_total = 0;
    int *h_counts = nullptr;  // pinned, nranks^2 ints
    // per-step halo over peer memory (CUDA IPC + NVLink stores); NCCL send/recv is the fallback
    int halo_mode = 0;  // 0 = auto (peer memory when it can be mapped), 1 = NCCL, 2 = peer memory or fail
    bool p2p_tried = false, p2p_ok = false;
    double4 *p2p_recv = nullptr;
    unsigned long long *p2p_sig = nullptr;
    int p2p_half = 0;
    double4 *peer_recv[P2P_MAX_RANKS] = {nullptr};
    unsigned long long *peer_sig[P2P_MAX_RANKS] = {nullptr};
    int p2p_dst_off[P2P_MAX_RANKS] = {0};
    unsigned long long halo_seq = 0;

    // CUDA graphs of NVE step batches (single GPU): the rebuild chain sits in a device-side conditional node
    struct StepGraph {
        int m = 0;
        int64_t nvt_total = 0;     // 0: NVE steps; > 0: NVT steps with this ramp length
        const void *f0 = nullptr;  // which of the two force buffers held F(t) when the batch was captured
        const void *x0 = nullptr;  // ... and which of the two position buffers held x(t) (fused steps alternate them)
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<StepGraph> graphs;
    std::vector<unsigned char> graph_sig;  // everything the captured launches baked in (pointers, box, grid, dt, ...)
    cudaStream_t graph_stream = nullptr;   // capture stream of the conditional bodies
    int use_graphs = 1;                    // option "cuda_graphs"
    bool capturing = false;
    int64_t n_graph_launches = 0;

    // stats
    int64_t n_steps = 0, n_launches = 0;
    int64_t max_nbr = 0, total_nbr = 0;
    int64_t n_builds_host = 0;

    // profiling
    bool profiling = false;
    struct Ev {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Ev> ev_pool;
    size_t ev_used = 0;
    double t_ms[PISB_K_COUNT] = {0};
    int64_t t_n[PISB_K_COUNT] = {0};
};

namespace {

int fail(pisb_t *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, PISB_ERR_CUDA, fmt("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                                              cudaGetErrorString(e_)));                            \
    } while (0)

#define TRY(expr)                      \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != PISB_OK) return rc_; \
    } while (0)

template <typename T>
int dev_reserve(pisb_t *h, DevBuf<T> &b, size_t n) {
    if (n <= b.cap) return PISB_OK;
    if (b.p) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaFree(b.p));
        h->device_bytes -= (int64_t)(b.cap * sizeof(T));
        b.p = nullptr;
        b.cap = 0;
    }
    CUDA_TRY(h, cudaMalloc((void **)&b.p, n * sizeof(T)));
    b.cap = n;
    h->device_bytes += (int64_t)(n * sizeof(T));
    return PISB_OK;
}

// Position buffers (xt / s_xt) carry one extra record IN FRONT of slot 0: {NaN, NaN, NaN, type 1}.  A list pad gathers it
// (index -1 after one IMNMX in the force loop); its squared distance is NaN and fails every range test.
int reserve_positions(pisb_t *h, DevBuf<double4> &b, size_t n) {
    if (n <= b.cap) return PISB_OK;
    if (b.p) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaFree(b.p - 1));
        h->device_bytes -= (int64_t)((b.cap + 1) * sizeof(double4));
        b.p = nullptr;
        b.cap = 0;
    }
    double4 *raw = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&raw, (n + 1) * sizeof(double4)));
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    double4 rec;
    rec.x = rec.y = rec.z = qnan;
    const long long type_one = 1;
    std::memcpy(&rec.w, &type_one, sizeof(double));
    CUDA_TRY(h, cudaMemcpy(raw, &rec, sizeof rec, cudaMemcpyHostToDevice));
    b.p = raw + 1;
    b.cap = n;
    h->device_bytes += (int64_t)((n + 1) * sizeof(double4));
    return PISB_OK;
}

void free_positions(pisb_t *h, DevBuf<double4> &b) {
    if (b.p) cudaFree(b.p - 1);
    if (b.cap) h->device_bytes -= (int64_t)((b.cap + 1) * sizeof(double4));
    b.p = nullptr;
    b.cap = 0;
}

// For buffers whose size follows a fluctuating count (migration / halo): grow with 50 % head-room so
// that the cudaFree + cudaMalloc pair (a device-wide synchronisation) happens O(log) times, not per rebuild.
template <typename T>
int dev_reserve_grow(pisb_t *h, DevBuf<T> &b, size_t n, size_t min_elems = 4096) {
    if (n <= b.cap) return PISB_OK;
    return dev_reserve(h, b, std::max(min_elems, n + n / 2));
}

template <typename T>
void dev_free(pisb_t *h, DevBuf<T> &b) {
    if (b.p) cudaFree(b.p);
    h->device_bytes -= (int64_t)(b.cap * sizeof(T));
    b.p = nullptr;
    b.cap = 0;
}

// ---- profiling -------------------------------------------------------------------------------
// One scope per launch (or launch group) of a kernel class: the launch counter, an NVTX range named after the class (the
// phases of a step show up as ranges in an Nsight timeline: the reference has no tracing), and with
// pisb_set_profiling an event pair for pisb_timings.
static const char *const kClassName[PISB_K_COUNT] = {"pisb:integrate", "pisb:bin", "pisb:sort", "pisb:build", "pisb:force",
                                                     "pisb:reduce", "pisb:halo", "pisb:copy"};
struct LaunchScope {
    pisb_t *h;
    int idx = -1;
    LaunchScope(pisb_t *h_, int cls) : h(h_) {
        h->n_launches++;
        nvtxRangePushA(kClassName[cls]);
        if (!h->profiling) return;
        if (h->ev_used == h->ev_pool.size()) {
            pisb_handle::Ev e;
            cudaEventCreate(&e.a);
            cudaEventCreate(&e.b);
            e.cls = cls;
            h->ev_pool.push_back(e);
        }
        idx = (int)h->ev_used++;
        h->ev_pool[idx].cls = cls;
        cudaEventRecord(h->ev_pool[idx].a, h->stream);
    }
    ~LaunchScope() {
        if (idx >= 0) cudaEventRecord(h->ev_pool[idx].b, h->stream);
        nvtxRangePop();
    }
};

void drain_events(pisb_t *h) {
    if (h->ev_used == 0) return;
    cudaStreamSynchronize(h->stream);
    for (size_t i = 0; i < h->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[i].a, h->ev_pool[i].b) == cudaSuccess) {
            h->t_ms[h->ev_pool[i].cls] += ms;
            h->t_n[h->ev_pool[i].cls] += 1;
        }
    }
    h->ev_used = 0;
}

inline int nblk(int n, int tpb) { return (n + tpb - 1) / tpb; }
// capped grid for the grid-stride kernels of the (conditionally executed) rebuild chain
inline int nblk_capped(int n, int tpb, int per_sm) { return std::max(1, std::min(nblk(n, tpb), 148 * per_sm)); }

int check_launch(pisb_t *h, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, PISB_ERR_CUDA, fmt("launch %s: %s", what, cudaGetErrorString(e)));
    return PISB_OK;
}

// ---- setup -----------------------------------------------------------------------------------
int build_pair_table(pisb_t *h) {
    const int nt = h->n_types;
    h->pairs.assign((size_t)nt * nt, PairDev{});
    h->max_rcut = 0.0;
    for (int k = 0; k < nt * nt; ++k) {
        PairDev &p = h->pairs[k];
        p.present = h->present[k] ? 1 : 0;
        if (!p.present) continue;
        const double e = h->eps[k], s = h->sigma[k], rc = h->rcut[k];
        if (h->max_rcut < rc) h->max_rcut = rc;  // PairPotentialManager::max_rcut (potential.rs:170-179)
        p.c4 = 4.0 * e;
        p.c24 = 24.0 * e;
        p.sig2 = s * s;
        p.t_rc = sqrt_threshold(rc);
        p.t_list = sqrt_threshold(rc + h->skin);
        p.t_lo = p.t_hi = p.t_rc;  // widened to the FP64 guard band once the box is known (setup_filter)
        if (h->shift) {
            // lennard_jones.rs:44-52, same expression order
            const double q = s / rc;
            const double ci2 = q * q;
            const double ca = (ci2 * ci2) * ci2;
            const double cr = ca * ca;
            p.ucut = 4.0 * e * (cr - ca);
        } else {
            p.ucut = 0.0;
        }
    }
    TRY(dev_reserve(h, h->table_d, (size_t)nt * nt));
    CUDA_TRY(h, cudaMemcpyAsync(h->table_d.p, h->pairs.data(), sizeof(PairDev) * nt * nt,
                                cudaMemcpyHostToDevice, h->stream));
    TRY(dev_reserve(h, h->mass_d, (size_t)nt));
    CUDA_TRY(h, cudaMemcpyAsync(h->mass_d.p, h->mass.data(), sizeof(double) * nt, cudaMemcpyHostToDevice,
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

// FP32 pre-filter constants (v2 kernels): guard band from the FP32 error bound
//   |r2f - r2| / r2 <= (17.4 L/r + 4) 2^-24   (positions rounded to f32 in [0, L], image shift by fl32(L))
// doubled for safety.  lo/hi are rounded outwards so the band can only grow.
float f32_below(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}
float f32_above(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}

bool v2_possible(const pisb_t *h) {
    return h->have_box && h->box.ortho && h->box.pbc[0] && h->box.pbc[1] && h->box.pbc[2];
}

// k_force_q -- four lanes per atom, integrator in the epilogue -- is the step kernel of systems too small to fill the GPU with
// one thread per atom (measured, profiles/r02_small_systems.jsonl: 32k atoms 44 against 55 us per step, 4000 atoms 30.0
// against 30.8; from 108k atoms up one thread per atom wins, 65 against 77 us, and at 4M atoms 1.18 against 1.67 ms per launch:
// 40 % more instructions for the same L1TEX wavefronts).  force_variant 7 forces it at any size.
bool quad_mode(const pisb_t *h) {
    return v2_possible(h) && (h->force_variant == 7 || (h->force_variant == 0 && h->n <= 75000));
}

int setup_filter(pisb_t *h) {
    if (!h->have_box) return PISB_OK;
    const int nt = h->n_types;
    // A box the FP32 / lean kernels cannot serve (triclinic: every NPT step after the first) runs on the reference-order
    // kernels, which read only the box-independent part of the table (c4, c24, sig2, t_rc, t_list, ucut): once it is on the
    // device there is nothing to recompute, upload or wait for when the box changes (a sync per NPT step).
    if (!v2_possible(h) && h->table_on_device) return PISB_OK;
    double lmax = 0.0;
    for (int d = 0; d < 3; ++d) {
        h->boxf.L[d] = (float)h->box.h[4 * d];
        h->boxf.invL[d] = (float)h->box.hinv[4 * d];
        h->boxf.pbc[d] = h->box.pbc[d];
        lmax = std::max(lmax, std::fabs(h->box.h[4 * d]));
    }
    h->boxf.margin = f32_above((h->max_rcut + 2.0 * h->skin) * 1.0001);
    h->pairsf.assign((size_t)nt * nt, PairF{-1.f, -1.f, -1.f, -1.f});
    for (int k = 0; k < nt * nt; ++k) {
        if (!h->pairs[k].present) continue;
        const double rc = h->rcut[k], rl = rc + h->skin;
        const double band_rc = (35.0 * lmax / rc + 8.0) * std::ldexp(1.0, -24);
        const double band_l = (35.0 * lmax / rl + 8.0) * std::ldexp(1.0, -24);
        PairF &f = h->pairsf[k];
        f.lo_rc = f32_below(h->pairs[k].t_rc * (1.0 - band_rc));
        f.hi_rc = f32_above(h->pairs[k].t_rc * (1.0 + band_rc));
        f.lo_list = f32_below(h->pairs[k].t_list * (1.0 - band_l));
        f.hi_list = f32_above(h->pairs[k].t_list * (1.0 + band_l));
    }
    // FP64 guard band of the lean force loop (pisb_kernels.cuh "Lean force loop"): the short-way r2 and the reference's
    // r2 = |h (h_inv d - round(h_inv d))|^2 differ by the rounding of the fractional round trip, a few ulp x L/|d|
    // relative; band = 2 x (4 L/rc + 32) ulp covers it with an order of magnitude to spare.  Rounded outwards.
    for (int k = 0; k < nt * nt; ++k) {
        PairDev &p = h->pairs[k];
        if (!p.present) continue;
        const double band = 2.0 * (4.0 * lmax / h->rcut[k] + 32.0) * std::ldexp(1.0, -52);
        p.t_lo = std::nextafter(p.t_rc * (1.0 - band), 0.0);
        p.t_hi = std::nextafter(p.t_rc * (1.0 + band), INFINITY);
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->table_d.p, h->pairs.data(), sizeof(PairDev) * nt * nt, cudaMemcpyHostToDevice, h->stream));
    TRY(dev_reserve(h, h->tablef_d, (size_t)nt * nt));
    h->table_on_device = true;
    if (!v2_possible(h)) {  // the FP32 pre-filter kernels are not selectable for this box
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return PISB_OK;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->tablef_d.p, h->pairsf.data(), sizeof(PairF) * nt * nt, cudaMemcpyHostToDevice,
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

// Cell grid: n_d = max(1, floor(|h col d| / (max_rcut + skin))) -- Atoms::divide_into_cells
// (src/atoms/neighbour_list.rs:45-57) with rcut := max_rcut + skin; capped so that the cell count
// stays O(N) for dilute systems (a coarser grid is still correct: cell edge >= list cutoff).
int setup_grid(pisb_t *h) {
    h->grid_ok = false;
    if (!h->have_box || !h->have_atoms) return PISB_OK;
    const double rc_list = h->max_rcut + h->skin;
    if (!(rc_list > 0.0)) return fail(h, PISB_ERR_INVALID, "no pair potential present (max_rcut = 0)");
    Grid g{};
    // Half-size cells (edge >= rc_list/2, 5^3 stencil) cut the candidate volume from 27 to 15.6 rc_list^3.
    // Only with the range-scanning v2 build and when every dimension has >= 5 such cells.
    // auto: half-size cells with the v3 build (1.81 vs 2.05 ms per build at 4M atoms; with the scalar v2 build the shorter
    // rows cost what the smaller candidate volume saved: profiles/r01_variants_build_v3.jsonl)
    int div = h->cell_div == 0 ? ((h->build_variant == 0 || h->build_variant == 3) && h->n_types == 1 ? 2 : 1) : h->cell_div;
    if (div < 1) div = 1;
    if (div > 2) div = 2;
    if (!(h->build_variant >= 2 || (h->build_variant == 0 && v2_possible(h)))) div = 1;
    const int64_t cell_cap = std::max<int64_t>(4 * (int64_t)h->n + 1024, 27);
    double prod = 1.0;
    // Extent that bounds the cell count of dimension d: the reference's column norm (divide_into_cells), and for a
    // tilted box also the perpendicular width 1 / |row d of h_inv| -- cells are slabs of the FRACTIONAL coordinate, so
    // it is the perpendicular width that must cover the list cutoff (the reference's column norm alone can make the
    // slabs thinner than rcut; a coarser grid is always still correct).
    auto extent = [&](int d) {
        const double *c = &h->box.h[d * 3];
        double len = std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
        if (!h->box.ortho) {
            const double *hi = h->box.hinv;
            const double rn = std::sqrt((hi[d] * hi[d] + hi[3 + d] * hi[3 + d]) + hi[6 + d] * hi[6 + d]);
            if (rn > 0.0) len = std::min(len, 1.0 / rn);
        }
        return len;
    };
    for (int d = 0; d < 3 && div > 1; ++d) {
        const double len = extent(d);
        const double nd = std::floor(len / (rc_list / div));
        if (nd < 2 * div + 1 || nd > 2048.0) div = 1;
        prod *= nd;
    }
    if (div > 1 && prod > (double)cell_cap) div = 1;  // dilute system: keep the cell count O(N)
    for (int d = 0; d < 3; ++d) {
        const double len = extent(d);
        double nd = std::floor(len / rc_list);
        if (!(nd >= 1.0)) nd = 1.0;
        if (div > 1) nd = std::floor(len / (rc_list / div));
        if (nd > 2048.0) nd = 2048.0;
        g.n[d] = (int)nd;
        if (h->box.pbc[d] && g.n[d] < 2 * div)
            return fail(h, PISB_ERR_INVALID,
                        fmt("box edge %d (%.6g) is shorter than 2 x (rcut + skin) = %.6g: the single-image "
                            "minimum-image convention of the reference (simulation_box.rs:17-27) is invalid",
                            d, len, 2.0 * rc_list));
    }
    while (div == 1 && (int64_t)g.n[0] * g.n[1] * g.n[2] > cell_cap) {
        int dmax = 0;
        for (int d = 1; d < 3; ++d)
            if (g.n[d] > g.n[dmax]) dmax = d;
        g.n[dmax] = std::max(3, g.n[dmax] * 3 / 4);
        if (g.n[0] <= 3 && g.n[1] <= 3 && g.n[2] <= 3) break;
    }
    for (int d = 0; d < 3; ++d) {
        // unique stencil offsets (the reference double-counts when n < 3: SURVEY appendix A.1)
        if (div > 1) g.lo[d] = -div, g.hi[d] = div;
        else if (g.n[d] == 1) g.lo[d] = 0, g.hi[d] = 0;
        else if (g.n[d] == 2) g.lo[d] = 0, g.hi[d] = 1;
        else g.lo[d] = -1, g.hi[d] = 1;
    }
    if (h->multi) {
        // decomposed dimensions: brick-local, non-periodic binning over brick + ghost shell
        if (!v2_possible(h)) return fail(h, PISB_ERR_INVALID, "multi-GPU mode needs an orthorhombic, fully periodic box");
        Decomp &dc = h->dc;
        dc.gw = rc_list * (1.0 + 1e-6);
        for (int d = 0; d < 3; ++d) {
            const double len = h->box.h[4 * d];
            const double w = len / dc.P[d];
            dc.lo[d] = dc.b[d] * w;
            dc.hi[d] = dc.b[d] == dc.P[d] - 1 ? len : (dc.b[d] + 1) * w;
            if (dc.P[d] == 1) continue;
            if (w < 2.0 * dc.gw)
                return fail(h, PISB_ERR_INVALID, fmt("brick width %.6g in dimension %d is below 2 x (rcut + skin) = %.6g", w, d, 2.0 * dc.gw));
            const double half = 0.5 * w + dc.gw + h->skin;
            int nd = (int)std::floor(2.0 * half / (rc_list / div));  // cell edge >= rc_list / div, stencil +-div
            if (nd < 2 * div + 1) nd = 2 * div + 1;
            g.n[d] = nd;
            g.local[d] = 1;
            g.center[d] = dc.lo[d] + 0.5 * w;
            g.half[d] = half;
            g.inv_edge[d] = nd / (2.0 * half);
            g.lo[d] = -div;
            g.hi[d] = div;
        }
    }
    g.ncell = g.n[0] * g.n[1] * g.n[2];
    h->grid = g;
    TRY(dev_reserve(h, h->cell_count, (size_t)g.ncell + 2));
    TRY(dev_reserve(h, h->cell_start, (size_t)g.ncell + 2));
    TRY(dev_reserve(h, h->tile_sum, (size_t)nblk(g.ncell + 1, SCAN_TILE) + 1));
    TRY(setup_filter(h));
    h->grid_ok = true;
    h->list_valid = false;
    return PISB_OK;
}

int estimate_kcap(pisb_t *h) {
    if (h->kcap_user > 0) return h->kcap_user;
    const double *m = h->box.h;
    const double vol = std::fabs(m[0] * (m[4] * m[8] - m[5] * m[7]) - m[3] * (m[1] * m[8] - m[2] * m[7]) +
                                 m[6] * (m[1] * m[5] - m[2] * m[4]));
    const double rl = h->max_rcut + h->skin;
    const double n_glob = h->multi ? (double)h->n_own * h->dc.nranks : (double)h->n;
    double k = vol > 0 ? n_glob / vol * 4.18879020478639 * rl * rl * rl : 64.0;
    k = 1.3 * k + 24.0;
    if (k > n_glob) k = std::max(n_glob, 1.0);
    if (k > 4096.0) k = 4096.0;
    return (int)std::ceil(k);
}

int reserve_atoms(pisb_t *h, int n) {
    const size_t cap = (size_t)n;
    TRY(reserve_positions(h, h->xt, cap));
    TRY(dev_reserve(h, h->xf, cap));
    TRY(dev_reserve(h, h->xf2, cap));
    TRY(dev_reserve(h, h->xp, ((cap + 1) / 2 + 32) * 8));
    TRY(reserve_positions(h, h->s_xt, cap));
    for (int d = 0; d < 3; ++d) {
        TRY(dev_reserve(h, h->v[d], cap));
        TRY(dev_reserve(h, h->f[d], cap));
        TRY(dev_reserve(h, h->g[d], cap));
        TRY(dev_reserve(h, h->xb[d], cap));
        TRY(dev_reserve(h, h->s_v[d], cap));
        TRY(dev_reserve(h, h->s_f[d], cap));
    }
    TRY(dev_reserve(h, h->id, cap));
    TRY(dev_reserve(h, h->s_id, cap));
    TRY(dev_reserve(h, h->slot_of_id, cap));
    TRY(dev_reserve(h, h->cell_of, cap));
    TRY(dev_reserve(h, h->order, cap));
    TRY(dev_reserve(h, h->nnbr, cap));
    // block_reduce_finalize: NQ x (blocks + groups) partial sums and 1 + groups tickets for the largest user --
    // k_force_vv (6 quantities, 128-thread blocks), k_npt_post (19 quantities, 256-thread blocks), k_force_split<8> (2
    // quantities, 8 lanes per atom below 75k atoms)
    auto need = [](size_t nq, size_t blocks) { return nq * (blocks + red_groups((unsigned int)blocks) + 2); };
    const size_t b_force = (size_t)nblk(n, TPB_FORCE), b_stream = (size_t)nblk(n, TPB), b_split = (size_t)nblk(std::min(n, 75000) * 8, TPB_FORCE);
    const size_t b_quad = ((size_t)n * 4 + TPB_Q - 1) / TPB_Q;  // k_force_q: four lanes per atom, 6 quantities
    TRY(dev_reserve(h, h->partials, std::max(std::max(need(6, b_force), need(6, b_quad)), std::max(need(19, b_stream), need(2, b_split)))));
    const size_t tickets = 2 + red_groups((unsigned int)std::max(std::max(b_force, b_quad), b_split));
    if (tickets > h->ticket_cap) {
        // zero-initialised once; the words reset themselves at the end of every reduction
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (h->ticket) cudaFree(h->ticket);
        h->ticket = nullptr;
        CUDA_TRY(h, cudaMalloc((void **)&h->ticket, sizeof(unsigned int) * tickets * 2));
        h->ticket_cap = tickets * 2;
        CUDA_TRY(h, cudaMemsetAsync(h->ticket, 0, sizeof(unsigned int) * h->ticket_cap, h->stream));
    }
    return PISB_OK;
}

int reserve_list(pisb_t *h) {
    h->npad = (std::max(h->n, h->ncap_atoms) + 31) / 32 * 32;
    h->kcap = (h->kcap + 3) / 4 * 4;  // whole K-tiles of 4
    TRY(dev_reserve(h, h->nbr, (size_t)h->kcap * (size_t)h->npad));
    return PISB_OK;
}

int reserve_thermo(pisb_t *h, size_t n) {
    TRY(dev_reserve(h, h->thermo_d, n));
    if (h->h_thermo_cap < n) {
        if (h->h_thermo) cudaFreeHost(h->h_thermo);
        h->h_thermo = nullptr;
        CUDA_TRY(h, cudaHostAlloc((void **)&h->h_thermo, sizeof(pisb_thermo) * n, cudaHostAllocDefault));
        h->h_thermo_cap = n;
    }
    return PISB_OK;
}

// Keep the list capacity AHEAD of the lists.  A build inside a batch cannot reallocate: if a row outgrows the capacity the
// batch goes on with a truncated row and the damage is only seen when the host reads the flags again (PISB_ERR_CAPACITY).
// So whenever the longest row seen since the last look (FLAG_MAXNBR, a running maximum) has come within 12 % of the
// capacity -- nothing truncated yet -- the capacity is raised now and the next call starts with a fresh build.  Density
// changes over many rebuilds, the host looks every batch (<= 4096 steps) or host-buffer step: an overflow needs the longest row to
// jump by more than 12 % between two looks.  Not on bricks (a rebuild is collective there) and not against a user-set capacity.
int grow_list_if_close(pisb_t *h) {
    if (h->kcap_user > 0 || h->multi) return PISB_OK;
    const int mx = h->h_flags[FLAG_MAXNBR];
    if (mx > h->kcap || (double)mx <= 0.88 * (double)h->kcap) return PISB_OK;
    h->kcap = (int)(mx * 1.3) + 8;
    h->list_valid = false;
    h->n_capacity_growths++;
    return reserve_list(h);
}

// ---- launch sequences ------------------------------------------------------------------------
template <typename F>
int dispatch_ortho(pisb_t *h, F &&f) {
    return h->box.ortho ? f(std::true_type{}) : f(std::false_type{});
}

// The rebuild chain: every kernel returns at once unless flags[FLAG_REBUILD] is set.
int launch_rebuild_chain(pisb_t *h) {
    const int n = h->n;
    const Grid g = h->grid;
    const int nb = g.ncell + (h->multi ? 1 : 0);  // + sentinel bucket for dead slots
    cudaStream_t st = h->stream;
    {
        LaunchScope ls(h, PISB_K_BIN);
        CUDA_TRY(h, cudaMemsetAsync(h->cell_count.p, 0, sizeof(int) * ((size_t)nb + 1), st));
        if (h->box.ortho)
            k_bin<true><<<nblk_capped(n, TPB, 8), TPB, 0, st>>>(n, h->xt.p, h->box, g, h->cell_of.p, h->cell_count.p, h->flags);
        else
            k_bin<false><<<nblk_capped(n, TPB, 8), TPB, 0, st>>>(n, h->xt.p, h->box, g, h->cell_of.p, h->cell_count.p, h->flags);
    }
    {
        LaunchScope ls(h, PISB_K_SORT);
        const int ntiles = nblk(nb, SCAN_TILE);
        k_scan_tiles<<<ntiles, SCAN_TPB, 0, st>>>(nb, h->cell_count.p, h->tile_sum.p, h->flags);
        k_scan_sums<<<1, SCAN_TPB, 0, st>>>(ntiles, h->tile_sum.p, h->flags);
        k_scan_apply<<<ntiles, SCAN_TPB, 0, st>>>(nb, n, h->cell_count.p, h->tile_sum.p, h->cell_start.p, h->flags);
        k_fill<<<nblk_capped(n, TPB, 8), TPB, 0, st>>>(n, g.ncell, h->cell_of.p, h->cell_start.p, h->cell_count.p, h->order.p, h->flags);
        k_sort_cells<<<nblk(g.ncell, TPB), TPB, 0, st>>>(g.ncell, h->cell_start.p, h->order.p, h->id.p, h->flags);
        PermArgs pa{n, h->order.p, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p,
                    h->id.p, h->s_xt.p, h->s_v[0].p, h->s_v[1].p, h->s_v[2].p, h->s_f[0].p, h->s_f[1].p,
                    h->s_f[2].p, h->s_id.p, h->flags};
        k_permute<<<nblk_capped(n, TPB, 8), TPB, 0, st>>>(pa);
        CopyBackArgs ca{n, h->s_xt.p, h->s_v[0].p, h->s_v[1].p, h->s_v[2].p, h->s_f[0].p, h->s_f[1].p, h->s_f[2].p,
                        h->s_id.p, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p,
                        h->id.p, h->multi ? nullptr : h->slot_of_id.p, h->xb[0].p, h->xb[1].p, h->xb[2].p, h->flags, h->xf.p,
                        h->box, h->xp.p};
        k_copy_back<<<nblk_capped(n, TPB, 8), TPB, 0, st>>>(ca);
        h->n_launches += 6;
    }
    {
        LaunchScope ls(h, PISB_K_BUILD);
        BuildArgs ba{n, h->npad, h->kcap, h->xt.p, h->cell_start.p, h->box, g, h->pairs[0], h->table_d.p,
                     h->n_types, h->nbr.p, h->nnbr.p, h->flags};
        const bool multi = h->n_types > 1;
        const int nb = nblk(n, TPB_FORCE);
        const bool v2 = h->build_variant >= 2 || (h->build_variant == 0 && v2_possible(h));
        if (v2) {
            if (!v2_possible(h)) return fail(h, PISB_ERR_INVALID, "build_variant 2/3 needs an orthorhombic, fully periodic box");
            Build2Args b2{n, h->npad, h->kcap, h->xt.p, h->xf.p, h->cell_start.p, h->box, h->boxf, g, h->pairs[0],
                          h->pairsf[0], h->table_d.p, h->tablef_d.p, h->n_types, h->nbr.p, h->nnbr.p, h->flags};
            if (h->build_variant == 2) {  // v2: scalar FP32 pre-filter (kept for A/B)
                if (multi) k_build_list_v2<true><<<nb, TPB_FORCE, 0, st>>>(b2);
                else k_build_list_v2<false><<<nb, TPB_FORCE, 0, st>>>(b2);
            } else {  // default: v3, packed FP32 pair records + bit-mask append for interior warps
                if (multi) k_build_list_v3<true><<<nb, TPB_FORCE, 0, st>>>(b2, h->xp.p, h->n_types <= B3_MAX_TYPES ? 1 : 0);
                else k_build_list_v3<false><<<nb, TPB_FORCE, 0, st>>>(b2, h->xp.p, 1);
            }
        } else if (h->box.ortho) {
            if (multi) k_build_list<true, true><<<nb, TPB_FORCE, 0, st>>>(ba);
            else k_build_list<true, false><<<nb, TPB_FORCE, 0, st>>>(ba);
        } else {
            if (multi) k_build_list<false, true><<<nb, TPB_FORCE, 0, st>>>(ba);
            else k_build_list<false, false><<<nb, TPB_FORCE, 0, st>>>(ba);
        }
        k_finish_rebuild<<<1, 1, 0, st>>>(h->flags);
        h->n_launches += 1;
    }
    return check_launch(h, "rebuild chain");
}

// f_out = (acc ? acc + LJ : LJ); thermo record gets pe and virial_pair.
int launch_force(pisb_t *h, double *const out[3], const double *const acc[3], pisb_thermo *rec, const int *skip_flag = nullptr) {
    LaunchScope ls(h, PISB_K_FORCE);
    ForceArgs fa{h->n, h->npad, h->xt.p, h->nbr.p, h->nnbr.p, h->box, h->pairs[0], h->table_d.p, h->n_types,
                 acc ? acc[0] : nullptr, acc ? acc[1] : nullptr, acc ? acc[2] : nullptr,
                 out[0], out[1], out[2], h->partials.p, h->ticket, rec};
    const bool multi = h->n_types > 1;
    const int nb = nblk(h->n, TPB_FORCE);
    cudaStream_t st = h->stream;
    const bool v2 = h->force_variant >= 2 || (h->force_variant == 0 && v2_possible(h));
    if (v2) {
        if (!v2_possible(h)) return fail(h, PISB_ERR_INVALID, "force_variant 2/3 needs an orthorhombic, fully periodic box");
        Force2Args f2{h->n, h->npad, h->xt.p, h->xf.p, h->nbr.p, h->nnbr.p, h->box, h->boxf, h->pairs[0], h->pairsf[0],
                      h->table_d.p, h->tablef_d.p, h->n_types, fa.ax, fa.ay, fa.az, out[0], out[1], out[2],
                      h->partials.p, h->ticket, rec, skip_flag, h->flags + h->xt_flag};
        if (quad_mode(h)) {  // four lanes per atom
            ForceVVArgs fv{};
            fv.f = f2;
            fv.flags = h->flags;
            const unsigned nbq = (unsigned)(((size_t)h->n * 4 + TPB_Q - 1) / TPB_Q);
            if (multi) {
                if (h->multi) k_force_q<true, false, false, true><<<nbq, TPB_Q, 0, st>>>(fv);
                else k_force_q<true, false, false, false><<<nbq, TPB_Q, 0, st>>>(fv);
            } else {
                if (h->multi) k_force_q<false, false, false, true><<<nbq, TPB_Q, 0, st>>>(fv);
                else k_force_q<false, false, false, false><<<nbq, TPB_Q, 0, st>>>(fv);
            }
            return check_launch(h, "k_force_q");
        }
        // only the default kernels honour skip_flag
        if (skip_flag && h->force_variant != 0 && h->force_variant != 3 && h->force_variant != 6)
            return fail(h, PISB_ERR_STATE, "speculative force launch needs force_variant 0, 3, 5 or 6");
        // systems that cannot fill the GPU with one thread per atom: S lanes per atom (force_variant 6 forces S = 8)
        const int split = h->force_variant == 6 ? 8 : 0;  // round-1 small-system kernel (whole K-tiles per lane), kept for A/B
        if (split) {
            const int nbs = nblk(h->n * split, TPB_FORCE);
            if (split == 8) {
                if (multi) k_force_split<true, 8><<<nbs, TPB_FORCE, 0, st>>>(f2);
                else k_force_split<false, 8><<<nbs, TPB_FORCE, 0, st>>>(f2);
            } else {
                if (multi) k_force_split<true, 4><<<nbs, TPB_FORCE, 0, st>>>(f2);
                else k_force_split<false, 4><<<nbs, TPB_FORCE, 0, st>>>(f2);
            }
        } else if (h->force_variant != 2) {  // auto = v3 (measured fastest: profiles/)
            if (multi) k_force_v3<true><<<nblk(h->n, TPB_VV), TPB_VV, 0, st>>>(f2);
            else k_force_v3<false><<<nblk(h->n, TPB_VV), TPB_VV, 0, st>>>(f2);
        } else {
            if (multi) k_force_v2<true><<<nb, TPB_FORCE, 0, st>>>(f2);
            else k_force_v2<false><<<nb, TPB_FORCE, 0, st>>>(f2);
        }
    } else if (h->box.ortho) {
        if (multi) k_force<true, true><<<nb, TPB_FORCE, 0, st>>>(fa);
        else k_force<true, false><<<nb, TPB_FORCE, 0, st>>>(fa);
    } else {
        if (multi) k_force<false, true><<<nb, TPB_FORCE, 0, st>>>(fa);
        else k_force<false, false><<<nb, TPB_FORCE, 0, st>>>(fa);
    }
    return check_launch(h, "k_force");
}

void swap_position_buffers(pisb_t *h) {
    std::swap(h->xt, h->s_xt);
    std::swap(h->xf, h->xf2);
    std::swap(h->xt_flag, h->s_xt_flag);
}

int launch_vv(pisb_t *h, bool kick, bool drift, double dt, pisb_thermo *rec, const double *vscale = nullptr,
              bool out_of_place = false) {
    LaunchScope ls(h, PISB_K_INTEGRATE);
    const double hs = 0.5 * h->skin;
    VVArgs a{h->n, h->xt.p, h->xf.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p,
             h->g[0].p, h->g[1].p, h->g[2].p, h->xb[0].p, h->xb[1].p, h->xb[2].p, h->mass_d.p, h->box,
             dt, dt * dt, h->skin_half2_override >= 0.0 ? h->skin_half2_override : hs * hs,
             (h->skin > 0.0 && h->skin_half2_override != 0.0) ? 0 : 1, h->flags, h->partials.p, h->ticket, rec, vscale,
             out_of_place ? h->s_xt.p : h->xt.p, out_of_place ? h->xf2.p : h->xf.p, out_of_place ? h->s_xt_flag : h->xt_flag};
    const int nb = nblk(h->n, TPB);
    cudaStream_t st = h->stream;
    const bool o = h->box.ortho != 0;
#define VV(K, D)                                                   \
    do {                                                           \
        if (o) k_vv<K, D, true><<<nb, TPB, 0, st>>>(a);            \
        else k_vv<K, D, false><<<nb, TPB, 0, st>>>(a);             \
    } while (0)
    if (kick && drift) VV(true, true);
    else if (kick) VV(true, false);
    else VV(false, true);
#undef VV
    if (out_of_place) swap_position_buffers(h);
    return check_launch(h, "k_vv");
}

// NVE batches of systems the default force kernel serves (orthorhombic, fully periodic, enough atoms for one thread per
// atom) run one k_force_vv launch per step instead of k_force_v3 + k_vv.
bool fused_step_possible(const pisb_t *h, bool multi_path = false) {
    if (!h->fuse_vv || h->multi != multi_path || !v2_possible(h)) return false;
    if (quad_mode(h)) return true;
    return (h->force_variant == 0 || h->force_variant == 3) && (h->force_variant == 3 || h->n > 75000);
}

// F(t+dt) -> g, kick with F(t) = f, [drift into the other position buffer], then f <-> g (and the position buffers).
// With a skip_flag the launch is speculative (multi-GPU) and the caller rotates the buffers once it knows the launch ran.
int launch_force_vv(pisb_t *h, bool drift, double dt, pisb_thermo *rec, const int *skip_flag = nullptr) {
    {
        LaunchScope ls(h, PISB_K_FORCE);
        const double hs = 0.5 * h->skin;
        ForceVVArgs fv{Force2Args{h->n, h->npad, h->xt.p, h->xf.p, h->nbr.p, h->nnbr.p, h->box, h->boxf, h->pairs[0], h->pairsf[0],
                                  h->table_d.p, h->tablef_d.p, h->n_types, nullptr, nullptr, nullptr, h->g[0].p, h->g[1].p, h->g[2].p,
                                  h->partials.p, h->ticket, rec, skip_flag, h->flags + h->xt_flag},
                       h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->xb[0].p, h->xb[1].p, h->xb[2].p,
                       h->mass_d.p, h->s_xt.p, h->xf2.p, dt, dt * dt,
                       h->skin_half2_override >= 0.0 ? h->skin_half2_override : hs * hs,
                       (h->skin > 0.0 && h->skin_half2_override != 0.0) ? 0 : 1, h->flags, h->s_xt_flag};
        const bool multi = h->n_types > 1;
        const int nbv = nblk(h->n, TPB_VV);
        cudaStream_t st = h->stream;
        if (quad_mode(h)) {
            const unsigned nbq = (unsigned)(((size_t)h->n * 4 + TPB_Q - 1) / TPB_Q);
#define FQ(T, D)                                                                  \
    do {                                                                          \
        if (h->multi) k_force_q<T, true, D, true><<<nbq, TPB_Q, 0, st>>>(fv);     \
        else k_force_q<T, true, D, false><<<nbq, TPB_Q, 0, st>>>(fv);             \
    } while (0)
            if (drift) {
                if (multi) FQ(true, true);
                else FQ(false, true);
            } else {
                if (multi) FQ(true, false);
                else FQ(false, false);
            }
#undef FQ
        } else {
#define FVV(T, D)                                                              \
    do {                                                                       \
        if (h->multi) k_force_vv<T, D, true><<<nbv, TPB_VV, 0, st>>>(fv);      \
        else k_force_vv<T, D, false><<<nbv, TPB_VV, 0, st>>>(fv);              \
    } while (0)
        if (drift) {
            if (multi) FVV(true, true);
            else FVV(false, true);
        } else {
            if (multi) FVV(true, false);
            else FVV(false, false);
        }
#undef FVV
        }
    }
    if (!skip_flag) {
        for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
        if (drift) swap_position_buffers(h);
    }
    return check_launch(h, "k_force_vv");
}

int read_flags(pisb_t *h) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->flags, sizeof(int) * FLAG_COUNT, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

int set_flag(pisb_t *h, int which, int value) {
    CUDA_TRY(h, cudaMemcpyAsync(h->flags + which, &value, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return PISB_OK;
}

// Make sure a valid list exists for the CURRENT positions (host-synchronous; grows capacity).
int ensure_list(pisb_t *h) {
    if (!h->grid_ok) TRY(setup_grid(h));
    if (!h->grid_ok) return fail(h, PISB_ERR_STATE, "set_box and upload must precede this call");
    // Nothing has touched the positions since the list was last confirmed: no chain launch, and above all no flag read-back.
    // That 32-byte device-to-host copy queues behind an asynchronous dump frame on the copy engine -- every batch of the
    // host's run loop started 1.75 ms late at 4M atoms (tools/diag_batches.py: 18.5 against 16.8 ms per 10-step batch).
    if (h->list_valid && h->list_checked && h->kcap > 0) return PISB_OK;
    if (h->kcap == 0) {
        h->kcap = estimate_kcap(h);
        TRY(reserve_list(h));
    }
    for (int attempt = 0; attempt < 6; ++attempt) {
        if (!h->list_valid) TRY(set_flag(h, FLAG_REBUILD, 1));
        TRY(set_flag(h, FLAG_MAXNBR, 0));
        TRY(launch_rebuild_chain(h));
        TRY(read_flags(h));
        const int mx = h->h_flags[FLAG_MAXNBR];
        if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
            h->n_builds_host = h->h_flags[FLAG_NBUILDS];
            h->max_nbr = mx;
        }
        if (mx <= h->kcap) {
            h->list_valid = true;
            h->list_checked = true;
            return PISB_OK;
        }
        if (h->kcap_user > 0 && mx > h->kcap_user)
            return fail(h, PISB_ERR_CAPACITY, fmt("neighbour list needs %d slots per atom, list_capacity is %d", mx, h->kcap_user));
        h->kcap = (int)(mx * 1.25) + 8;
        h->list_valid = false;
        TRY(reserve_list(h));
    }
    return fail(h, PISB_ERR_CAPACITY, "neighbour-list capacity did not converge");
}

// Stage a host AoS array on the device (pageable or pinned source).
int stage_in(pisb_t *h, DevBuf<double> &dst, const double *src, size_t count) {
    TRY(dev_reserve(h, dst, count));
    CUDA_TRY(h, cudaMemcpyAsync(dst.p, src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
    return PISB_OK;
}

// Type pairs the atoms populate but the table does not hold (their pairs are skipped, like the reference's `None => continue`,
// lennard_jones.rs:216-222).  Host-side, only when types arrive and the table has a hole at all.
void count_missing_pairs(pisb_t *h, int64_t n, const int32_t *types) {
    h->missing_type_pairs = 0;
    const int nt = h->n_types;
    bool hole = false;
    for (int a = 0; a < nt; ++a)  // the table is read at (min, max): the upper triangle is what counts
        for (int b = a; b < nt; ++b) hole = hole || !h->pairs[(size_t)a * nt + b].present;
    if (!hole || !types) return;
    std::vector<int64_t> pop((size_t)nt, 0);
    for (int64_t i = 0; i < n; ++i)
        if (types[i] >= 1 && types[i] <= nt) pop[(size_t)types[i] - 1]++;
    for (int a = 0; a < nt; ++a)
        for (int b = a; b < nt; ++b)
            if (!h->pairs[(size_t)a * nt + b].present && pop[a] > 0 && pop[b] > (a == b ? 1 : 0)) h->missing_type_pairs++;
}

int do_upload(pisb_t *h, int64_t n64, const double *pos, const double *vel, const double *frc,
              const int32_t *types) {
    if (n64 <= 0 || n64 > 2000000000LL) return fail(h, PISB_ERR_INVALID, "atom count out of range");
    if (!pos) return fail(h, PISB_ERR_INVALID, "pos is NULL");
    const int n = (int)n64;
    const bool same_set = h->have_atoms && h->n == n && h->list_valid && types == nullptr;
    if (!h->have_atoms && !types) return fail(h, PISB_ERR_INVALID, "types is NULL on first upload");
    if (h->have_atoms && h->n != n && !types) return fail(h, PISB_ERR_INVALID, "types is NULL but the atom count changed");
    if (!same_set) {
        if (h->n != n) {
            h->kcap = 0;
            h->grid_ok = false;
        }
        h->n = n;
        h->have_atoms = true;
        h->list_valid = false;
        TRY(reserve_atoms(h, n));
    }
    h->list_checked = false;  // new positions: the device decides again (k_check_displacement / the rebuild chain)
    const size_t n3 = (size_t)3 * n;
    TRY(stage_in(h, h->st_pos, pos, n3));
    if (vel) TRY(stage_in(h, h->st_vel, vel, n3));
    if (frc) TRY(stage_in(h, h->st_frc, frc, n3));
    if (types) {
        TRY(dev_reserve(h, h->st_types, (size_t)n));
        CUDA_TRY(h, cudaMemcpyAsync(h->st_types.p, types, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
        count_missing_pairs(h, n, types);
    }
    {
        LaunchScope ls(h, PISB_K_COPY);
        LoadArgs la{n, h->st_pos.p, vel ? h->st_vel.p : nullptr, frc ? h->st_frc.p : nullptr,
                    h->st_types.p, same_set ? h->slot_of_id.p : nullptr, h->xt.p,
                    h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p, h->n_types, h->flags,
                    (same_set && h->have_box) ? h->xf.p : nullptr, h->box, nullptr, h->have_box ? 1 : 0, h->xt_flag};
        TRY(set_flag(h, h->xt_flag, 0));
        k_load_aos<<<nblk(n, TPB), TPB, 0, h->stream>>>(la);
        TRY(check_launch(h, "k_load_aos"));
        if (same_set && h->have_box) {
            // keep the list if no atom moved more than skin/2 from its build position
            const double hs = 0.5 * h->skin;
            if (h->skin > 0.0) {
                if (h->box.ortho)
                    k_check_displacement<true><<<nblk(n, TPB), TPB, 0, h->stream>>>(n, h->xt.p, h->xb[0].p, h->xb[1].p, h->xb[2].p, h->box, hs * hs, h->flags);
                else
                    k_check_displacement<false><<<nblk(n, TPB), TPB, 0, h->stream>>>(n, h->xt.p, h->xb[0].p, h->xb[1].p, h->xb[2].p, h->box, hs * hs, h->flags);
                h->n_launches++;
            } else {
                TRY(set_flag(h, FLAG_REBUILD, 1));
            }
        }
    }
    h->forces_current = false;
    return PISB_OK;
}

int do_download(pisb_t *h, double *pos, double *vel, double *frc) {
    if (!h->have_atoms) return fail(h, PISB_ERR_STATE, "download before upload");
    const int n = h->n;
    const size_t n3 = (size_t)3 * n;
    if (pos) TRY(dev_reserve(h, h->st_pos, n3));
    if (vel) TRY(dev_reserve(h, h->st_vel, n3));
    if (frc) TRY(dev_reserve(h, h->st_frc, n3));
    {
        LaunchScope ls(h, PISB_K_COPY);
        StoreArgs sa{n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p,
                     pos ? h->st_pos.p : nullptr, vel ? h->st_vel.p : nullptr, frc ? h->st_frc.p : nullptr};
        k_store_aos<<<nblk(n, TPB), TPB, 0, h->stream>>>(sa);
        TRY(check_launch(h, "k_store_aos"));
    }
    if (pos) CUDA_TRY(h, cudaMemcpyAsync(pos, h->st_pos.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    if (vel) CUDA_TRY(h, cudaMemcpyAsync(vel, h->st_vel.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    if (frc) CUDA_TRY(h, cudaMemcpyAsync(frc, h->st_frc.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

// pisb_download_begin: snapshot into buffers of its own (the staging arrays belong to uploads and host-buffer steps),
// copies on copy_stream behind an event, so the main stream is free for the next batch at once.
int do_download_begin(pisb_t *h, double *pos, double *vel, double *frc) {
    if (!h->have_atoms) return fail(h, PISB_ERR_STATE, "download before upload");
    if (h->multi) return fail(h, PISB_ERR_STATE, "pisb_download_begin is a single-GPU entry point (use pisb_download_owned)");
    if (h->dl_pending) return fail(h, PISB_ERR_STATE, "a download is already in flight: call pisb_download_end first");
    const int n = h->n;
    const size_t n3 = (size_t)3 * n;
    if (pos) TRY(dev_reserve(h, h->dl_pos, n3));
    if (vel) TRY(dev_reserve(h, h->dl_vel, n3));
    if (frc) TRY(dev_reserve(h, h->dl_frc, n3));
    {
        LaunchScope ls(h, PISB_K_COPY);
        StoreArgs sa{n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p,
                     pos ? h->dl_pos.p : nullptr, vel ? h->dl_vel.p : nullptr, frc ? h->dl_frc.p : nullptr};
        k_store_aos<<<nblk(n, TPB), TPB, 0, h->stream>>>(sa);
        TRY(check_launch(h, "k_store_aos"));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_pos, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_pos, 0));
    h->dl_pending = true;
    if (pos) CUDA_TRY(h, cudaMemcpyAsync(pos, h->dl_pos.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->copy_stream));
    if (vel) CUDA_TRY(h, cudaMemcpyAsync(vel, h->dl_vel.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->copy_stream));
    if (frc) CUDA_TRY(h, cudaMemcpyAsync(frc, h->dl_frc.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->copy_stream));
    return PISB_OK;
}

// Issue queued pieces of an owned-atom frame worth up to `budget` bytes on dl_stream (0 = everything that is left).
int pump_download(pisb_t *h, size_t budget) {
    size_t sent = 0;
    while (h->dl_next < h->dl_queue.size() && (budget == 0 || sent < budget)) {
        const pisb_handle::DlPiece &p = h->dl_queue[h->dl_next++];
        CUDA_TRY(h, cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDeviceToHost, h->dl_stream));
        sent += p.bytes;
    }
    return PISB_OK;
}

int do_download_end(pisb_t *h) {
    if (!h->dl_pending) return PISB_OK;
    if (h->multi && h->dl_stream) TRY(pump_download(h, 0));
    h->dl_queue.clear();
    h->dl_next = 0;
    h->dl_pending = false;
    CUDA_TRY(h, cudaStreamSynchronize(h->multi && h->dl_stream ? h->dl_stream : h->copy_stream));
    return PISB_OK;
}

int check_bad_type(pisb_t *h) {
    if (h->h_flags[FLAG_BADTYPE] != 0) {
        int o = h->h_flags[FLAG_BADTYPE];
        set_flag(h, FLAG_BADTYPE, 0);
        return fail(h, PISB_ERR_INVALID, fmt("atom %d has a type outside 1..%d", o, h->n_types));
    }
    return PISB_OK;
}

int do_compute_multi(pisb_t *h, int accumulate, double *pe);
int do_step_nve_multi(pisb_t *h, double dt, int64_t nsteps, pisb_thermo *out);

int do_compute(pisb_t *h, int accumulate, double *pe) {
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "set_box and upload must precede compute");
    if (h->multi) return do_compute_multi(h, accumulate, pe);
    TRY(ensure_list(h));
    TRY(check_bad_type(h));
    TRY(reserve_thermo(h, 2));
    double *out[3] = {h->f[0].p, h->f[1].p, h->f[2].p};
    const double *acc[3] = {h->f[0].p, h->f[1].p, h->f[2].p};
    TRY(launch_force(h, out, accumulate ? acc : nullptr, h->thermo_d.p));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (pe) *pe = h->h_thermo[0].pe;
    h->forces_current = true;
    return PISB_OK;
}

// ---- NVE step batches ----------------------------------------------------------------------------------------
// Enqueue cnt steps writing records rec[0..cnt): drift, [rebuild], force, (kick+drift, [rebuild], force) x (cnt-1), kick.
// The kick of step k is fused with the drift of step k+1.  While a graph is being captured the rebuild chain goes
// into an IF node whose condition a one-thread kernel copies from flags[REBUILD]; otherwise the chain is launched and
// every kernel of it returns at once unless the flag is set.
int capture_conditional_rebuild(pisb_t *h);

int enqueue_nve_steps(pisb_t *h, double dt, int64_t cnt, pisb_thermo *rec0) {
    if (fused_step_possible(h)) {
        // drift (into the other position buffer), [rebuild], (force+kick+drift, [rebuild]) x (cnt-1), force+kick:
        // cnt position-buffer swaps and cnt force-buffer swaps, so an even batch ends on the buffers it started from
        TRY(launch_vv(h, false, true, dt, nullptr, nullptr, true));
        for (int64_t s = 0; s < cnt; ++s) {
            if (h->capturing) TRY(capture_conditional_rebuild(h));
            else TRY(launch_rebuild_chain(h));
            TRY(launch_force_vv(h, s + 1 < cnt, dt, rec0 + s));
        }
        return PISB_OK;
    }
    for (int64_t s = 0; s < cnt; ++s) {
        pisb_thermo *rec = rec0 + s;
        if (s == 0) TRY(launch_vv(h, false, true, dt, nullptr));
        else TRY(launch_vv(h, true, true, dt, rec - 1));
        if (h->capturing) TRY(capture_conditional_rebuild(h));
        else TRY(launch_rebuild_chain(h));
        double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
        TRY(launch_force(h, outp, nullptr, rec));
        for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
    }
    return launch_vv(h, true, false, dt, rec0 + (cnt - 1));
}

__global__ void k_set_condition(cudaGraphConditionalHandle handle, const int *flags) {
    cudaGraphSetConditional(handle, flags[FLAG_REBUILD] != 0 ? 1u : 0u);
}

int capture_conditional_rebuild(pisb_t *h) {
    cudaStreamCaptureStatus status;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    CUDA_TRY(h, cudaStreamGetCaptureInfo(h->stream, &status, nullptr, &graph, &deps, &ndeps));
    if (status != cudaStreamCaptureStatusActive) return fail(h, PISB_ERR_STATE, "conditional rebuild outside a capture");
    cudaGraphConditionalHandle cond;
    CUDA_TRY(h, cudaGraphConditionalHandleCreate(&cond, graph, 0, cudaGraphCondAssignDefault));
    k_set_condition<<<1, 1, 0, h->stream>>>(cond, h->flags);
    TRY(check_launch(h, "k_set_condition"));
    CUDA_TRY(h, cudaStreamGetCaptureInfo(h->stream, &status, nullptr, &graph, &deps, &ndeps));
    cudaGraphNodeParams np{};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeIf;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    CUDA_TRY(h, cudaGraphAddNode(&node, graph, deps, ndeps, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    CUDA_TRY(h, cudaStreamBeginCaptureToGraph(h->graph_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    cudaStream_t main_stream = h->stream;
    h->stream = h->graph_stream;  // the chain's launches land in the body graph
    const int rc = launch_rebuild_chain(h);
    h->stream = main_stream;
    cudaGraph_t ended = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->graph_stream, &ended);
    if (rc != PISB_OK) return rc;
    if (e != cudaSuccess) return fail(h, PISB_ERR_CUDA, fmt("capturing the rebuild chain: %s", cudaGetErrorString(e)));
    CUDA_TRY(h, cudaStreamUpdateCaptureDependencies(h->stream, &node, 1, cudaStreamSetCaptureDependencies));
    return PISB_OK;
}

void drop_graphs(pisb_t *h) {
    for (auto &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// Everything a captured launch sequence bakes in.  Compared byte-wise before every replay.
void graph_signature(pisb_t *h, double dt, std::vector<unsigned char> &sig) {
    sig.clear();
    auto put = [&](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        sig.insert(sig.end(), b, b + n);
    };
    // f and g trade places every step: the pair is part of the signature, the current assignment is the graph's key
    const void *fa = std::min<const void *>(h->f[0].p, h->g[0].p), *fb = std::max<const void *>(h->f[0].p, h->g[0].p);
    // so do the two position buffers (and FP32 shadows) of fused steps
    const void *ptrs[] = {std::min<const void *>(h->xt.p, h->s_xt.p), std::max<const void *>(h->xt.p, h->s_xt.p),
                          std::min<const void *>(h->xf.p, h->xf2.p), std::max<const void *>(h->xf.p, h->xf2.p), h->xp.p, h->tablef_d.p, h->v[0].p, h->v[1].p, h->v[2].p, fa, fb,
                          std::min<const void *>(h->f[1].p, h->g[1].p), std::max<const void *>(h->f[1].p, h->g[1].p),
                          std::min<const void *>(h->f[2].p, h->g[2].p), std::max<const void *>(h->f[2].p, h->g[2].p), h->xb[0].p, h->xb[1].p, h->xb[2].p, h->s_v[0].p,
                          h->s_v[1].p, h->s_v[2].p, h->s_f[0].p, h->s_f[1].p, h->s_f[2].p, h->id.p, h->s_id.p,
                          h->slot_of_id.p, h->cell_of.p, h->order.p, h->nnbr.p, h->nbr.p, h->cell_count.p, h->cell_start.p,
                          h->tile_sum.p, h->mass_d.p, h->partials.p, h->table_d.p, h->thermo_d.p, h->flags, h->ticket, h->nhc_d.p,
                          h->nhc_energy_d.p};
    put(ptrs, sizeof ptrs);
    const int ints[] = {h->n, h->npad, h->kcap, h->n_types, h->force_variant, h->build_variant, h->cell_div, h->fuse_vv,
                        quad_mode(h) ? 1 : 0};
    put(ints, sizeof ints);
    const double dbl[] = {dt, h->skin, h->skin_half2_override, h->max_rcut};
    put(dbl, sizeof dbl);
    for (int k = 0; k < 9; ++k) put(&h->box.h[k], sizeof(double)), put(&h->box.hinv[k], sizeof(double));
    put(h->box.pbc, sizeof h->box.pbc);
    put(&h->box.ortho, sizeof(int));
    put(h->grid.n, sizeof h->grid.n), put(h->grid.lo, sizeof h->grid.lo), put(h->grid.hi, sizeof h->grid.hi);
    put(&h->boxf.margin, sizeof(float));
    for (const PairDev &pd : h->pairs) {
        const double v[] = {pd.c4, pd.c24, pd.sig2, pd.t_rc, pd.t_list, pd.ucut, (double)pd.present, pd.t_lo, pd.t_hi};
        put(v, sizeof v);
    }
}

// Executable graph of m NVE steps on records thermo_d[0..m) (m even: the f/g swap returns to the captured assignment).
int enqueue_nvt_steps(pisb_t *h, double dt, int64_t cnt, int64_t total_steps, pisb_thermo *rec0, double *energy0);

int capture_step_graph(pisb_t *h, double dt, int m, int64_t nvt_total) {
    if (!h->graph_stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
    const int64_t launches_before = h->n_launches;
    // nothing executes during a capture, but the enqueue functions rotate the force / position buffers as they go: a capture
    // that fails half way must leave the handle's buffer assignment exactly as it found it
    const DevBuf<double> f_before[3] = {h->f[0], h->f[1], h->f[2]}, g_before[3] = {h->g[0], h->g[1], h->g[2]};
    const DevBuf<double4> xt_before = h->xt, s_xt_before = h->s_xt;
    const DevBuf<float4> xf_before = h->xf, xf2_before = h->xf2;
    auto restore = [&]() {
        for (int d = 0; d < 3; ++d) h->f[d] = f_before[d], h->g[d] = g_before[d];
        h->xt = xt_before, h->s_xt = s_xt_before, h->xf = xf_before, h->xf2 = xf2_before;
    };
    CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
    const int rc = nvt_total > 0 ? enqueue_nvt_steps(h, dt, m, nvt_total, h->thermo_d.p, h->nhc_energy_d.p)
                                 : enqueue_nve_steps(h, dt, m, h->thermo_d.p);
    h->capturing = false;
    h->n_launches = launches_before;  // captured, not launched
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc != PISB_OK || e != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        restore();
        if (rc != PISB_OK) return rc;
        return fail(h, PISB_ERR_CUDA, fmt("cudaStreamEndCapture: %s", cudaGetErrorString(e)));
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(h, PISB_ERR_CUDA, fmt("cudaGraphInstantiate: %s", cudaGetErrorString(ei)));
    pisb_handle::StepGraph sg;
    sg.m = m;
    sg.nvt_total = nvt_total;
    sg.f0 = h->f[0].p;
    sg.x0 = h->xt.p;
    sg.exec = exec;
    h->graphs.push_back(sg);
    return PISB_OK;
}

constexpr int GRAPH_M = 8;  // steps per replay; batches are cut into pieces of 8, 4, 2 (and a classic single step)

// All piece sizes, for both assignments of the force buffers, are captured together the first time a state signature is
// seen, so the one-off capture + instantiation cost (~1 ms per captured step, ~30 ms in all) is paid in the first batch,
// not whenever a new remainder or an odd step count shows up.
int get_step_graph(pisb_t *h, double dt, int m, int64_t nvt_total, cudaGraphExec_t *out) {
    std::vector<unsigned char> sig;
    graph_signature(h, dt, sig);
    if (sig != h->graph_sig) {
        drop_graphs(h);
        h->graph_sig = sig;
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (auto &g : h->graphs)
            if (g.m == m && g.nvt_total == nvt_total && g.f0 == h->f[0].p && g.x0 == h->xt.p) {
                *out = g.exec;
                return PISB_OK;
            }
        if (pass == 0) {
            if (h->graphs.size() > 24) drop_graphs(h);
            // both assignments of the two force buffers (an odd classic step flips it), so that no later batch pays
            int rc = PISB_OK;
            for (int parity = 0; parity < 2 && rc == PISB_OK; ++parity) {
                for (int mm = GRAPH_M; mm >= 2 && rc == PISB_OK; mm /= 2) rc = capture_step_graph(h, dt, mm, nvt_total);
                // pointers only (nothing executes during a capture); fused steps flip the position buffers with the force buffers
                const bool fused = nvt_total == 0 && fused_step_possible(h);
                for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
                if (fused) swap_position_buffers(h);
                if (rc != PISB_OK && parity == 0) {  // leave the buffers as they were
                    for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
                    if (fused) swap_position_buffers(h);
                }
            }
            if (rc != PISB_OK) return rc;
        }
    }
    return fail(h, PISB_ERR_STATE, "no graph for this batch size");
}

int do_step_nve(pisb_t *h, double dt, int64_t nsteps, pisb_thermo *out) {
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "set_box and upload must precede step_nve");
    if (nsteps < 0) return fail(h, PISB_ERR_INVALID, "nsteps < 0");
    if (nsteps == 0) return PISB_OK;
    if (h->multi) return do_step_nve_multi(h, dt, nsteps, out);
    TRY(ensure_list(h));  // list for x(t); from here on the skin trigger decides on the device
    TRY(check_bad_type(h));
    const int64_t chunk_max = 4096;
    // Graph replay needs launch sequences without per-kernel timing events.
    const bool graphs = h->use_graphs && !h->profiling;
    int64_t done = 0;
    while (done < nsteps) {
        const int64_t m = std::min(chunk_max, nsteps - done);
        TRY(reserve_thermo(h, (size_t)chunk_max + 1));  // fixed size: the record pointer is baked into the graphs
        const int builds_before = (int)h->n_builds_host;
        int64_t off = 0, graph_steps = 0;
        while (off < m) {
            int64_t piece = m - off;
            cudaGraphExec_t exec = nullptr;
            if (graphs && h->use_graphs && piece >= 2) {
                const int gm = piece >= GRAPH_M ? GRAPH_M : (piece >= 4 ? 4 : 2);
                if (get_step_graph(h, dt, gm, 0, &exec) == PISB_OK) {
                    piece = gm;
                } else {  // e.g. a driver without conditional graph nodes: same kernels, classic launches from now on
                    h->use_graphs = 0;
                    exec = nullptr;
                    drop_graphs(h);
                    cudaGetLastError();
                }
            }
            if (exec) {
                CUDA_TRY(h, cudaGraphLaunch(exec, h->stream));
                h->n_graph_launches++;
                graph_steps += piece;
                // fixed nodes: piece condition setters + either one drift and piece k_force_vv launches, or piece+1 integrator
                // and piece force launches
                h->n_launches += (fused_step_possible(h) ? 2 : 3) * piece + 1;
            } else {
                TRY(enqueue_nve_steps(h, dt, piece, h->thermo_d.p));
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo + off, h->thermo_d.p, sizeof(pisb_thermo) * piece, cudaMemcpyDeviceToHost, h->stream));
            off += piece;
        }
        TRY(read_flags(h));
        if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
            h->n_builds_host = h->h_flags[FLAG_NBUILDS];
            h->max_nbr = h->h_flags[FLAG_MAXNBR];
        }
        // kernels of the conditional bodies that did run: 10 per build (an upper bound when classic pieces were mixed in)
        if (graph_steps > 0) h->n_launches += 10 * (int64_t)((int)h->n_builds_host - builds_before);
        if (h->h_flags[FLAG_MAXNBR] > h->kcap) {
            const int mx = h->h_flags[FLAG_MAXNBR];
            h->list_valid = false;
            if (h->kcap_user == 0) {
                h->kcap = (int)(mx * 1.25) + 8;
                TRY(reserve_list(h));
            }
            return fail(h, PISB_ERR_CAPACITY,
                        fmt("a neighbour list overflowed during the batch (needed %d slots per atom); capacity was "
                            "grown -- re-upload the state and repeat the call", mx));
        }
        if (out) std::memcpy(out + done, h->h_thermo, sizeof(pisb_thermo) * m);
        done += m;
        h->n_steps += m;
        TRY(grow_list_if_close(h));
        if (!h->list_valid && done < nsteps) TRY(ensure_list(h));
    }
    h->forces_current = true;
    return PISB_OK;
}



// NVT batch: verlet_step_nvt_nhc x nsteps with the chain on the device (see k_nhc_half).  The kinetic energy a first half
// step starts from and the ramp's step index are carried in NhcDev, so a step sequence is the same launch list for every
// step and can be replayed as a graph (same 8 / 4 / 2 pieces and conditional rebuild node as the NVE batches).
int enqueue_nvt_steps(pisb_t *h, double dt, int64_t cnt, int64_t total_steps, pisb_thermo *rec0, double *energy0) {
    for (int64_t s = 0; s < cnt; ++s) {
        pisb_thermo *rec = rec0 + s;
        k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, nullptr, (long long)h->n, dt, 0, -1, 1, nullptr);
        TRY(launch_vv(h, false, true, dt, nullptr, &h->nhc_d.p->scale));
        if (h->capturing) TRY(capture_conditional_rebuild(h));
        else TRY(launch_rebuild_chain(h));
        double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
        TRY(launch_force(h, outp, nullptr, rec));
        for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
        TRY(launch_vv(h, true, false, dt, rec, &h->nhc_d.p->scale));
        k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, &rec->ke, (long long)h->n, dt, 1, -1, (long long)total_steps, energy0 + s);
        h->n_launches += 2;
    }
    return check_launch(h, "nvt step");
}

int do_step_nvt_multi(pisb_t *h, double dt, int64_t nsteps, pisb_nhc *chain, int64_t first_step, int64_t total_steps,
                      pisb_thermo *out, double *nhc_energy);
int wait_published(pisb_t *h, volatile int *pub, int seq);

int do_step_nvt(pisb_t *h, double dt, int64_t nsteps, pisb_nhc *chain, int64_t first_step, int64_t total_steps,
                pisb_thermo *out, double *nhc_energy) {
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "set_box and upload must precede step_nvt_nhc");
    if (h->multi) return do_step_nvt_multi(h, dt, nsteps, chain, first_step, total_steps, out, nhc_energy);
    if (!chain || chain->chain_size != 3) return fail(h, PISB_ERR_INVALID, "chain must be a 3-link pisb_nhc (pisb_nhc_init)");
    if (nsteps < 0 || total_steps <= 0) return fail(h, PISB_ERR_INVALID, "bad step counts");
    if (nsteps == 0) return PISB_OK;
    TRY(ensure_list(h));
    TRY(check_bad_type(h));
    const int64_t chunk_max = 4096;
    TRY(dev_reserve(h, h->nhc_d, 1));
    TRY(dev_reserve(h, h->nhc_energy_d, (size_t)chunk_max));
    TRY(reserve_thermo(h, (size_t)chunk_max + 1));
    {
        // KE of the state entering the batch (potential.rs:41), reduced on the device, seeds the carried value
        pisb_thermo *ke0 = h->thermo_d.p + chunk_max;
        LaunchScope ls(h, PISB_K_REDUCE);
        k_observe<<<nblk(h->n, TPB), TPB, 0, h->stream>>>(h->n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p,
                                                          h->f[2].p, h->mass_d.p, h->partials.p, h->ticket, ke0);
        TRY(check_launch(h, "k_observe"));
        // the record reaches the host through a publish kernel, not a copy: a 32-byte device-to-host copy would queue behind
        // an asynchronous dump frame on the copy engine and start every batch of a dumping NVT run late (see ensure_list)
        const int seq = ++h->pub_seq, words = (int)(sizeof(pisb_thermo) / sizeof(int));
        k_publish_words<<<1, 32, 0, h->stream>>>(reinterpret_cast<const int *>(ke0), h->h_pub, words, seq);
        TRY(wait_published(h, h->h_pub + words, seq));
        std::memcpy(&h->h_thermo[0], h->h_pub, sizeof(pisb_thermo));
    }
    NhcDev init{};
    init.c = *chain;
    init.scale = 1.0;
    init.ke_last = h->h_thermo[0].ke;
    init.step_index = first_step;
    CUDA_TRY(h, cudaMemcpyAsync(h->nhc_d.p, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    const bool graphs = h->use_graphs && !h->profiling;
    int64_t done = 0;
    if (h->h_energy_cap < (size_t)chunk_max) {
        if (h->h_energy) cudaFreeHost(h->h_energy);
        h->h_energy = nullptr, h->h_energy_cap = 0;
        CUDA_TRY(h, cudaHostAlloc((void **)&h->h_energy, sizeof(double) * chunk_max, cudaHostAllocDefault));
        h->h_energy_cap = (size_t)chunk_max;
    }
    double *he = h->h_energy;
    while (done < nsteps) {
        const int64_t m = std::min(chunk_max, nsteps - done);
        const int builds_before = (int)h->n_builds_host;
        int64_t off = 0, graph_steps = 0;
        while (off < m) {
            int64_t piece = m - off;
            cudaGraphExec_t exec = nullptr;
            if (graphs && h->use_graphs && piece >= 2) {
                const int gm = piece >= GRAPH_M ? GRAPH_M : (piece >= 4 ? 4 : 2);
                if (get_step_graph(h, dt, gm, total_steps, &exec) == PISB_OK) {
                    piece = gm;
                } else {
                    h->use_graphs = 0;
                    exec = nullptr;
                    drop_graphs(h);
                    cudaGetLastError();
                }
            }
            if (exec) {
                CUDA_TRY(h, cudaGraphLaunch(exec, h->stream));
                h->n_graph_launches++;
                graph_steps += piece;
                h->n_launches += 6 * piece;  // 2 chain + 2 integrator + condition + force per step
            } else {
                TRY(enqueue_nvt_steps(h, dt, piece, total_steps, h->thermo_d.p, h->nhc_energy_d.p));
            }
            CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo + off, h->thermo_d.p, sizeof(pisb_thermo) * piece, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaMemcpyAsync(he + off, h->nhc_energy_d.p, sizeof(double) * piece, cudaMemcpyDeviceToHost, h->stream));
            off += piece;
        }
        TRY(read_flags(h));
        if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
            h->n_builds_host = h->h_flags[FLAG_NBUILDS];
            h->max_nbr = h->h_flags[FLAG_MAXNBR];
        }
        if (graph_steps > 0) h->n_launches += 10 * (int64_t)((int)h->n_builds_host - builds_before);
        if (h->h_flags[FLAG_MAXNBR] > h->kcap) {
            h->list_valid = false;
            return fail(h, PISB_ERR_CAPACITY, "a neighbour list overflowed during the batch; re-upload the state and repeat the call");
        }
        if (out) std::memcpy(out + done, h->h_thermo, sizeof(pisb_thermo) * m);
        if (nhc_energy) std::memcpy(nhc_energy + done, he, sizeof(double) * m);
        done += m;
        h->n_steps += m;
        TRY(grow_list_if_close(h));
        if (!h->list_valid && done < nsteps) TRY(ensure_list(h));
    }
    NhcDev fin{};
    CUDA_TRY(h, cudaMemcpy(&fin, h->nhc_d.p, sizeof fin, cudaMemcpyDeviceToHost));
    *chain = fin.c;
    h->forces_current = true;
    return PISB_OK;
}

// NPT batch: verlet_step_npt_mtk x nsteps (potential.rs:112-135).  The barostat (nine momenta) is advanced on the host from
// the tensors the device reduces; one small device->host read per step (see pisb_npt.cuh).
void box_from_matrices(BoxDev &b, const double *h9, const double *hinv9) {
    bool ortho = true;
    for (int k = 0; k < 9; ++k) {
        b.h[k] = h9[k];
        b.hinv[k] = hinv9[k];
        if (k % 4 != 0 && (h9[k] != 0.0 || hinv9[k] != 0.0)) ortho = false;
    }
    b.ortho = ortho ? 1 : 0;
}

// MTKBarostat::delta_momentum (npt.rs:45-50) from the reduced tensors t[0..8] = V V^T, t[9..17] = X F^T
void mtk_delta_momentum(const pisb_mtk &m, const double *t, double volume, double dt, double *out) {
    double p[9];
    const double f = volume * 0.5 * dt;
    for (int k = 0; k < 9; ++k) p[k] = ((t[k] + t[9 + k]) / volume - m.target_pressure[k]) * f;
    m3::symmetrize(p, out);
}

// MTKBarostat::scale (npt.rs:52-58)
void mtk_scale(const pisb_mtk &m, double dt, bool velocity_scaling, double *out) {
    double e[9];
    for (int k = 0; k < 9; ++k) e[k] = m.momentum[k] / m.w;
    m3::symmetrize(e, e);
    const double factor = velocity_scaling ? -0.5 : 1.0;
    for (int k = 0; k < 9; ++k) e[k] = e[k] * factor * dt;
    m3::expm(e, out);
}

int launch_npt_post(pisb_t *h, const double *scale9, pisb_thermo *rec) {
    LaunchScope ls(h, PISB_K_REDUCE);
    NptPostArgs a{};
    a.n = h->n;
    a.xt = h->xt.p;
    a.vx = h->v[0].p, a.vy = h->v[1].p, a.vz = h->v[2].p;
    a.fx = h->f[0].p, a.fy = h->f[1].p, a.fz = h->f[2].p;
    a.mass = h->mass_d.p;
    a.apply_scale = scale9 ? 1 : 0;
    if (scale9) std::memcpy(a.scale.m, scale9, sizeof a.scale.m);
    a.partials = h->partials.p;
    a.ticket = h->ticket;
    a.tensors = h->npt_tensors_d.p;
    a.thermo = rec;
    k_npt_post<<<nblk(h->n, TPB), TPB, 0, h->stream>>>(a);
    return check_launch(h, "k_npt_post");
}

int do_step_npt(pisb_t *h, double dt, int64_t nsteps, pisb_mtk *baro, pisb_nhc *chain, int64_t first_step,
                int64_t total_steps, pisb_thermo *out, double *ext_energy, double *h9_trace) {
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "set_box and upload must precede step_npt_mtk");
    if (h->multi) return fail(h, PISB_ERR_STATE, "pisb_step_npt_mtk is a single-GPU entry point");
    if (!baro || !(baro->w > 0.0)) return fail(h, PISB_ERR_INVALID, "barostat must come from pisb_mtk_init (w > 0)");
    if (!chain || chain->chain_size != 3) return fail(h, PISB_ERR_INVALID, "chain must be a 3-link pisb_nhc (pisb_nhc_init)");
    if (nsteps < 0 || total_steps <= 0) return fail(h, PISB_ERR_INVALID, "bad step counts");
    if (nsteps == 0) return PISB_OK;
    TRY(ensure_list(h));
    h->list_checked = false;  // the box moves under the list from here on; the batch keeps its own account of it
    TRY(check_bad_type(h));
    TRY(dev_reserve(h, h->nhc_d, 1));
    TRY(dev_reserve(h, h->nhc_energy_d, 1));
    TRY(dev_reserve(h, h->npt_tensors_d, 19));
    TRY(dev_reserve(h, h->partials, (size_t)19 * (nblk(h->n, TPB) + red_groups((unsigned int)nblk(h->n, TPB)) + 2)));
    TRY(reserve_thermo(h, 4));
    if (!h->h_npt) CUDA_TRY(h, cudaHostAlloc((void **)&h->h_npt, sizeof(double) * 32, cudaHostAllocDefault));
    NhcDev init{};
    init.c = *chain;
    init.scale = 1.0;
    CUDA_TRY(h, cudaMemcpyAsync(h->nhc_d.p, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    pisb_thermo *rec = h->thermo_d.p, *ke0 = h->thermo_d.p + 1;
    const double rc_list = h->max_rcut + h->skin;
    double h_build[9], tensors[19];
    // A list that predates this call was built in this very box: pisb_set_box invalidates the list on any change, and
    // an NPT batch that ends with a strained list invalidates it too (below), so F = I here.
    std::memcpy(h_build, h->box.h, sizeof h_build);
    TRY(launch_npt_post(h, nullptr, nullptr));  // tensors of the state entering the batch
    CUDA_TRY(h, cudaMemcpyAsync(h->h_npt, h->npt_tensors_d.p, sizeof(double) * 19, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    std::memcpy(tensors, h->h_npt, sizeof tensors);
    int rc = PISB_OK;
    for (int64_t s = 0; s < nsteps && rc == PISB_OK; ++s) {
        rc = [&]() -> int {
            double dm[9], scale[9], scale_h[9], h_new[9], hinv_new[9];
            double volume = std::fabs(m3::det(h->box.h));
            mtk_delta_momentum(*baro, tensors, volume, dt, dm);
            for (int k = 0; k < 9; ++k) baro->momentum[k] += dm[k];
            mtk_scale(*baro, dt, true, scale);
            mtk_scale(*baro, dt, false, scale_h);
            m3::mul(scale_h, h->box.h, h_new);  // scale_box: h = scale * h; h_inv = try_inverse (transformations.rs:8-13)
            if (!m3::inverse(h_new, hinv_new)) return fail(h, PISB_ERR_INVALID, "Box matrix should be invertible");
            for (int k = 0; k < 9; ++k)
                if (!std::isfinite(h_new[k]) || !std::isfinite(hinv_new[k])) return fail(h, PISB_ERR_INVALID, "the barostat drove the box to a non-finite matrix");
            {
                LaunchScope ls(h, PISB_K_INTEGRATE);
                NptPreArgs a{};
                a.n = h->n;
                a.xt = h->xt.p, a.xf = h->xf.p;
                a.vx = h->v[0].p, a.vy = h->v[1].p, a.vz = h->v[2].p;
                a.xbx = h->xb[0].p, a.xby = h->xb[1].p, a.xbz = h->xb[2].p;
                a.mass = h->mass_d.p;
                std::memcpy(a.scale.m, scale, sizeof scale);
                std::memcpy(a.hinv_old.m, h->box.hinv, sizeof scale);
                std::memcpy(a.h_new.m, h_new, sizeof scale);
                a.partials = h->partials.p, a.ticket = h->ticket, a.ke_out = ke0;
                k_npt_pre<<<nblk(h->n, TPB), TPB, 0, h->stream>>>(a);
                TRY(check_launch(h, "k_npt_pre"));
            }
            // new box: grid (cells are slabs of the fractional coordinate, so the grid only changes when a cell count does)
            {
                const Grid old = h->grid;
                const bool lv = h->list_valid;
                box_from_matrices(h->box, h_new, hinv_new);
                TRY(setup_grid(h));
                const bool same = old.n[0] == h->grid.n[0] && old.n[1] == h->grid.n[1] && old.n[2] == h->grid.n[2];
                h->list_valid = lv && same;
                if (!h->list_valid) {
                    TRY(ensure_list(h));
                    std::memcpy(h_build, h->box.h, sizeof h_build);
                }
            }
            // skin budget left after the affine strain since the last build: a pair outside the list had
            // |r_build| > rc + skin and now |r| >= sigma_min(F) |r_build| - 2 u_max, F = h h_build^-1
            {
                double hbi[9], F[9];
                if (!m3::inverse(h_build, hbi)) return fail(h, PISB_ERR_INVALID, "singular build box");
                m3::mul(h->box.h, hbi, F);
                double e2 = 0.0;
                for (int k = 0; k < 9; ++k) {
                    const double d = F[k] - (k % 4 == 0 ? 1.0 : 0.0);
                    e2 += d * d;
                }
                const double room = (1.0 - std::sqrt(e2)) * rc_list - h->max_rcut;
                h->skin_half2_override = room > 0.0 ? 0.25 * room * room * (1.0 - 1e-12) : 0.0;
            }
            k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, &ke0->ke, (long long)h->n, dt, 0, 0, 1, nullptr);
            TRY(launch_vv(h, false, true, dt, nullptr, &h->nhc_d.p->scale));
            TRY(launch_rebuild_chain(h));
            double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
            TRY(launch_force(h, outp, nullptr, rec));
            for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
            TRY(launch_vv(h, true, false, dt, rec, &h->nhc_d.p->scale));
            k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, &rec->ke, (long long)h->n, dt, 1, (long long)(first_step + s),
                                                (long long)total_steps, h->nhc_energy_d.p);
            h->n_launches += 2;
            TRY(launch_npt_post(h, scale, rec));  // v = S v (potential.rs:129) + the tensors of the new state
            CUDA_TRY(h, cudaMemcpyAsync(h->h_npt, h->npt_tensors_d.p, sizeof(double) * 19, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaMemcpyAsync(h->h_npt + 19, h->nhc_energy_d.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, rec, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
            TRY(read_flags(h));
            if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
                h->n_builds_host = h->h_flags[FLAG_NBUILDS];
                h->max_nbr = h->h_flags[FLAG_MAXNBR];
                std::memcpy(h_build, h->box.h, sizeof h_build);
            }
            if (h->h_flags[FLAG_MAXNBR] > h->kcap) {
                h->list_valid = false;
                return fail(h, PISB_ERR_CAPACITY, "a neighbour list overflowed during the batch; re-upload the state and repeat the call");
            }
            TRY(grow_list_if_close(h));  // a compressing box makes the rows longer: raise the capacity before one overflows (the next step rebuilds)
            std::memcpy(tensors, h->h_npt, sizeof tensors);
            volume = std::fabs(m3::det(h->box.h));
            mtk_delta_momentum(*baro, tensors, volume, dt, dm);
            for (int k = 0; k < 9; ++k) baro->momentum[k] += dm[k];
            if (out) out[s] = h->h_thermo[0];
            if (ext_energy) {
                // nhc KE + PE (device) + mtk.kinetic_energy() + mtk.potential_energy(h) (npt.rs:60-66)
                double mt[9], pt[9], pr[9];
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 3; ++r) {
                        mt[c * 3 + r] = baro->momentum[r * 3 + c];
                        pt[c * 3 + r] = baro->target_pressure[r * 3 + c];
                    }
                m3::mul(baro->momentum, mt, pr);
                const double mke = ((pr[0] + pr[4]) + pr[8]) / (2.0 * baro->w);
                m3::mul(pt, h->box.h, pr);
                const double mpe = (pr[0] + pr[4]) + pr[8];
                ext_energy[s] = ((h->h_npt[19] + mke) + mpe);
            }
            if (h9_trace) std::memcpy(h9_trace + 9 * s, h->box.h, sizeof(double) * 9);
            h->n_steps += 1;
            return PISB_OK;
        }();
    }
    h->skin_half2_override = -1.0;
    if (std::memcmp(h_build, h->box.h, sizeof h_build) != 0) h->list_valid = false;  // the next batch starts from a fresh list
    if (rc != PISB_OK) return rc;
    NhcDev fin{};
    CUDA_TRY(h, cudaMemcpy(&fin, h->nhc_d.p, sizeof fin, cudaMemcpyDeviceToHost));
    *chain = fin.c;
    h->forces_current = true;
    return PISB_OK;
}

// ================================================================================================
// multi-GPU host orchestration (see pisb_multi.cuh for the scheme)
// ================================================================================================
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

// NCCL is loaded lazily (dlopen) so that single-GPU users do not need it; under torch the already
// loaded libnccl.so.2 is picked up.
int load_nccl(pisb_t *h) {
    if (g_nccl.lib) return PISB_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(h, PISB_ERR_COMM, fmt("cannot load libnccl.so.2: %s", dlerror()));
#define SYM(field, name)                                                             \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                    \
    if (!g_nccl.field) return fail(h, PISB_ERR_COMM, "libnccl is missing symbol " name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return PISB_OK;
}

#define NCCL_TRY(h, expr)                                                                              \
    do {                                                                                               \
        ncclResult_t r_ = (expr);                                                                      \
        if (r_ != ncclSuccess)                                                                         \
            return fail(h, PISB_ERR_COMM, fmt("%s failed at line %d: %s", #expr, __LINE__, g_nccl.GetErrorString(r_))); \
    } while (0)

// all-gather `cnt` (nranks ints per rank) and bring the nranks^2 table to the host (synchronous)
int wait_published(pisb_t *h, volatile int *pub, int seq);

int gather_counts(pisb_t *h) {
    const int R = h->dc.nranks;
    NCCL_TRY(h, g_nccl.AllGather(h->m_cnt.p, h->m_allcnt.p, R, ncclInt, h->comm, h->stream));
    // to the host without a copy engine (a D2H copy would queue behind an asynchronous dump frame: 8.6 ms at 8 GPUs)
    const int seq = ++h->pub_seq;
    k_publish_words<<<1, 32, 0, h->stream>>>(h->m_allcnt.p, h->h_counts, R * R, seq);
    return wait_published(h, h->h_counts + R * R, seq);
}

// grouped pairwise exchange of `rec` doubles per item with every other rank
int exchange(pisb_t *h, const double *sbuf, const std::vector<int> &scnt, const std::vector<int> &soff, double *rbuf,
             const std::vector<int> &rcnt, const std::vector<int> &roff, int rec) {
    const int R = h->dc.nranks, me = h->dc.rank;
    LaunchScope ls(h, PISB_K_HALO);
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int r = 0; r < R; ++r) {
        if (r == me) continue;
        if (scnt[r] > 0)
            NCCL_TRY(h, g_nccl.Send(sbuf + (size_t)soff[r] * rec, (size_t)scnt[r] * rec, ncclDouble, r, h->comm, h->stream));
        if (rcnt[r] > 0)
            NCCL_TRY(h, g_nccl.Recv(rbuf + (size_t)roff[r] * rec, (size_t)rcnt[r] * rec, ncclDouble, r, h->comm, h->stream));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    return PISB_OK;
}

double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int p2p_setup(pisb_t *h);

// Atom migration + ghost selection + cell sort + list build.  Host-synchronous (counts cross PCIe).
int multi_rebuild(pisb_t *h) {
    const bool trace = std::getenv("PISB_TRACE") != nullptr;
    double tw[8];
    tw[0] = wall_now();
    if (!h->grid_ok) TRY(setup_grid(h));
    if (!h->grid_ok) return fail(h, PISB_ERR_STATE, "set_box and upload must precede this call");
    const int R = h->dc.nranks, me = h->dc.rank;
    cudaStream_t st = h->stream;
    TRY(dev_reserve(h, h->m_cnt, (size_t)R));
    TRY(dev_reserve(h, h->m_allcnt, (size_t)R * R));
    TRY(dev_reserve(h, h->m_off, (size_t)R));
    TRY(dev_reserve(h, h->m_dest, (size_t)h->ncap_atoms));
    TRY(dev_reserve(h, h->m_pig, (size_t)h->ncap_atoms));
    TRY(dev_reserve(h, h->newslot, (size_t)h->ncap_atoms));
    TRY(dev_reserve(h, h->ghost_slot, (size_t)h->ncap_atoms));
    std::vector<int> scnt(R), rcnt(R), soff(R), roff(R);

    // ---- 1. migration: owned atoms whose wrapped position left the brick go to their new owner ----
    const int n_slots0 = h->n;  // live + (after this step) dead slots
    {
        LaunchScope ls(h, PISB_K_HALO);
        CUDA_TRY(h, cudaMemsetAsync(h->m_cnt.p, 0, sizeof(int) * R, st));
        k_mig_count<<<nblk(std::max(n_slots0, 1), TPB), TPB, 0, st>>>(n_slots0, h->xt.p, h->box, h->dc, h->m_dest.p, h->m_pig.p, h->m_cnt.p);
        TRY(check_launch(h, "k_mig_count"));
    }
    tw[1] = wall_now();
    TRY(gather_counts(h));
    tw[2] = wall_now();
    int stot = 0, rtot = 0;
    for (int r = 0; r < R; ++r) {
        scnt[r] = r == me ? 0 : h->h_counts[me * R + r];
        rcnt[r] = r == me ? 0 : h->h_counts[r * R + me];
        soff[r] = stot;
        roff[r] = rtot;
        stot += scnt[r];
        rtot += rcnt[r];
    }
    const int n_own = h->n_own - stot + rtot;
    if (n_slots0 + rtot > h->ncap_atoms)
        return fail(h, PISB_ERR_CAPACITY, fmt("rank %d: %d slots + %d arrivals exceed the local capacity %d", me, n_slots0, rtot, h->ncap_atoms));
    TRY(dev_reserve_grow(h, h->mig_send, (size_t)std::max(stot, 1) * MIG_REC));
    TRY(dev_reserve_grow(h, h->mig_recv, (size_t)std::max(rtot, 1) * MIG_REC));
    CUDA_TRY(h, cudaMemcpyAsync(h->m_off.p, soff.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
    {
        LaunchScope ls(h, PISB_K_HALO);
        MigArgs ma{n_slots0, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p,
                   h->m_dest.p, h->m_pig.p, h->m_off.p, h->mig_send.p};
        k_mig_pack<<<nblk(std::max(n_slots0, 1), TPB), TPB, 0, st>>>(ma);
        TRY(check_launch(h, "k_mig_pack"));
    }
    if (stot + rtot > 0 || R > 1) TRY(exchange(h, h->mig_send.p, scnt, soff, h->mig_recv.p, rcnt, roff, MIG_REC));
    if (rtot > 0) {
        LaunchScope ls(h, PISB_K_HALO);
        MigUnpackArgs u2{rtot, n_slots0, h->mig_recv.p, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p,
                         h->f[0].p, h->f[1].p, h->f[2].p, h->id.p};
        k_mig_unpack<<<nblk(rtot, TPB), TPB, 0, st>>>(u2);
        TRY(check_launch(h, "k_mig_unpack"));
    }
    h->n_own = n_own;
    const int n_slots1 = n_slots0 + rtot;  // owned atoms (old + arrived) live somewhere in [0, n_slots1)

    tw[3] = wall_now();
    // ---- 2. ghosts: owned atoms within gw of a brick face go to the rank(s) across ----
    {
        LaunchScope ls(h, PISB_K_HALO);
        CUDA_TRY(h, cudaMemsetAsync(h->m_cnt.p, 0, sizeof(int) * R, st));
        k_ghost_select<false><<<nblk(std::max(n_slots1, 1), TPB), TPB, 0, st>>>(n_slots1, h->xt.p, h->dc, h->m_cnt.p, nullptr, nullptr);
        TRY(check_launch(h, "k_ghost_select"));
    }
    TRY(gather_counts(h));
    h->gs_cnt.assign(R, 0);
    h->gr_cnt.assign(R, 0);
    h->gs_off.assign(R, 0);
    h->gr_off.assign(R, 0);
    int gs = 0, gr = 0;
    for (int r = 0; r < R; ++r) {
        h->gs_cnt[r] = r == me ? 0 : h->h_counts[me * R + r];
        h->gr_cnt[r] = r == me ? 0 : h->h_counts[r * R + me];
        h->gs_off[r] = gs;
        h->gr_off[r] = gr;
        gs += h->gs_cnt[r];
        gr += h->gr_cnt[r];
    }
    h->send_total = gs;
    h->n_ghost = gr;
    if (!h->p2p_tried) TRY(p2p_setup(h));
    if (h->p2p_ok) {
        if (gr > h->p2p_half) return fail(h, PISB_ERR_CAPACITY, "ghost count exceeds the peer-memory receive buffer");
        for (int r = 0; r < R; ++r) {  // where my block starts inside peer r's receive layout (its gr_off[me])
            int off = 0;
            for (int q = 0; q < me; ++q)
                if (q != r) off += h->h_counts[q * R + r];
            h->p2p_dst_off[r] = off;
        }
    }
    if (n_slots1 + gr > h->ncap_atoms)
        return fail(h, PISB_ERR_CAPACITY, fmt("rank %d needs %d local slots (owned + ghost + stale), capacity %d", me, n_slots1 + gr, h->ncap_atoms));
    TRY(dev_reserve_grow(h, h->send_idx, (size_t)std::max(gs, 1)));
    TRY(dev_reserve_grow(h, h->halo_send, (size_t)std::max(gs, 1)));
    TRY(dev_reserve_grow(h, h->halo_recv, (size_t)std::max(gr, 1)));
    TRY(dev_reserve_grow(h, h->mig_send, (size_t)std::max(gs, 1) * GHOST_REC));
    TRY(dev_reserve_grow(h, h->mig_recv, (size_t)std::max(gr, 1) * GHOST_REC));
    CUDA_TRY(h, cudaMemcpyAsync(h->m_off.p, h->gs_off.data(), sizeof(int) * R, cudaMemcpyHostToDevice, st));
    {
        LaunchScope ls(h, PISB_K_HALO);
        CUDA_TRY(h, cudaMemsetAsync(h->m_cnt.p, 0, sizeof(int) * R, st));
        k_ghost_select<true><<<nblk(std::max(n_slots1, 1), TPB), TPB, 0, st>>>(n_slots1, h->xt.p, h->dc, h->m_cnt.p, h->m_off.p, h->send_idx.p);
        if (gs > 0) k_ghost_pack_full<<<nblk(gs, TPB), TPB, 0, st>>>(gs, h->send_idx.p, h->xt.p, h->id.p, h->mig_send.p);
        TRY(check_launch(h, "k_ghost_pack_full"));
        h->n_launches += 1;
    }
    TRY(exchange(h, h->mig_send.p, h->gs_cnt, h->gs_off, h->mig_recv.p, h->gr_cnt, h->gr_off, GHOST_REC));
    if (gr > 0) {
        LaunchScope ls(h, PISB_K_HALO);
        k_ghost_append<<<nblk(gr, TPB), TPB, 0, st>>>(gr, n_slots1, h->mig_recv.p, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p,
                                                     h->f[0].p, h->f[1].p, h->f[2].p, h->id.p);
        TRY(check_launch(h, "k_ghost_append"));
    }
    h->n = n_slots1 + gr;  // slots entering the sort; the dead ones are dropped by it
    tw[4] = wall_now();

    // ---- 3. cell sort + list build over owned + ghost atoms; remap the halo index lists ----
    if (h->kcap == 0) {
        h->kcap = estimate_kcap(h);
        TRY(reserve_list(h));
    }
    for (int attempt = 0; attempt < 6; ++attempt) {
        TRY(set_flag(h, FLAG_REBUILD, 1));
        TRY(set_flag(h, FLAG_MAXNBR, 0));
        TRY(launch_rebuild_chain(h));
        {
            LaunchScope ls(h, PISB_K_HALO);
            k_inverse_perm<<<nblk(h->n, TPB), TPB, 0, st>>>(h->n, h->order.p, h->newslot.p);
            if (gs > 0) k_remap<<<nblk(gs, TPB), TPB, 0, st>>>(gs, h->send_idx.p, h->newslot.p);
            if (gr > 0) {
                if (attempt == 0) k_ghost_slots<<<nblk(gr, TPB), TPB, 0, st>>>(gr, n_slots1, h->newslot.p, h->ghost_slot.p);
                else k_remap<<<nblk(gr, TPB), TPB, 0, st>>>(gr, h->ghost_slot.p, h->newslot.p);
            }
            TRY(check_launch(h, "halo remap"));
            h->n_launches += 2;
        }
        h->n = n_own + gr;  // live slots (cell-sorted) after the sort compacted the dead ones away
        tw[5] = wall_now();
        {
            const int seq = ++h->pub_seq;
            k_publish_flags<<<1, 32, 0, st>>>(h->flags, h->h_flags, seq, 0);
            TRY(wait_published(h, h->h_flags + FLAG_COUNT, seq));
        }
        tw[6] = wall_now();
        const int mx = h->h_flags[FLAG_MAXNBR];
        h->n_builds_host = h->h_flags[FLAG_NBUILDS];
        h->max_nbr = mx;
        if (mx <= h->kcap) {
            h->list_valid = true;
            if (trace)
                fprintf(stderr, "[pisb rank %d] rebuild wall (ms): count-launch %.3f gather1 %.3f mig %.3f ghosts %.3f chain-launch %.3f chain-sync %.3f | leave %d arrive %d ghosts %d\n",
                        me, 1e3 * (tw[1] - tw[0]), 1e3 * (tw[2] - tw[1]), 1e3 * (tw[3] - tw[2]), 1e3 * (tw[4] - tw[3]),
                        1e3 * (tw[5] - tw[4]), 1e3 * (tw[6] - tw[5]), stot, rtot, gr);
            return PISB_OK;
        }
        if (h->kcap_user > 0) return fail(h, PISB_ERR_CAPACITY, fmt("neighbour list needs %d slots per atom, list_capacity is %d", mx, h->kcap_user));
        h->kcap = (int)(mx * 1.25) + 8;
        TRY(reserve_list(h));
    }
    return fail(h, PISB_ERR_CAPACITY, "neighbour-list capacity did not converge");
}


// ---- peer-memory halo: map every peer's receive buffer and signal words through CUDA IPC ----------------
struct P2PBlob {
    cudaIpcMemHandle_t recv, sig;
};

void p2p_teardown(pisb_t *h) {
    for (int r = 0; r < P2P_MAX_RANKS; ++r) {
        if (h->peer_recv[r]) cudaIpcCloseMemHandle(h->peer_recv[r]);
        if (h->peer_sig[r]) cudaIpcCloseMemHandle(h->peer_sig[r]);
        h->peer_recv[r] = nullptr;
        h->peer_sig[r] = nullptr;
    }
    if (h->p2p_recv) cudaFree(h->p2p_recv);
    if (h->p2p_sig) cudaFree(h->p2p_sig);
    h->p2p_recv = nullptr;
    h->p2p_sig = nullptr;
    h->p2p_ok = false;
}

int p2p_setup(pisb_t *h) {
    h->p2p_tried = true;
    h->p2p_ok = false;
    const int R = h->dc.nranks, me = h->dc.rank;
    if (R < 2 || R > P2P_MAX_RANKS || h->halo_mode == 1) return PISB_OK;
    cudaStream_t st = h->stream;
    // a common half size: the largest local capacity of any rank
    int *d_tmp = h->m_cnt.p;
    CUDA_TRY(h, cudaMemcpyAsync(d_tmp, &h->ncap_atoms, sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(h, g_nccl.AllReduce(d_tmp, d_tmp, 1, ncclInt, ncclMax, h->comm, st));
    int half = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&half, d_tmp, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    int ok = 1;
    if (cudaMalloc((void **)&h->p2p_recv, sizeof(double4) * 2 * (size_t)half) != cudaSuccess ||
        cudaMalloc((void **)&h->p2p_sig, sizeof(unsigned long long) * 2 * P2P_MAX_RANKS) != cudaSuccess)
        ok = 0;
    P2PBlob mine{};
    if (ok) {
        cudaMemset(h->p2p_sig, 0, sizeof(unsigned long long) * 2 * P2P_MAX_RANKS);
        if (cudaIpcGetMemHandle(&mine.recv, h->p2p_recv) != cudaSuccess || cudaIpcGetMemHandle(&mine.sig, h->p2p_sig) != cudaSuccess) ok = 0;
    }
    cudaGetLastError();
    // all-gather the handles (bytes) over the NCCL communicator
    DevBuf<char> d_blob;
    TRY(dev_reserve(h, d_blob, sizeof(P2PBlob) * (size_t)(R + 1)));
    CUDA_TRY(h, cudaMemcpyAsync(d_blob.p + sizeof(P2PBlob) * R, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
    NCCL_TRY(h, g_nccl.AllGather(d_blob.p + sizeof(P2PBlob) * R, d_blob.p, sizeof(P2PBlob), ncclChar, h->comm, st));
    std::vector<P2PBlob> all(R);
    CUDA_TRY(h, cudaMemcpyAsync(all.data(), d_blob.p, sizeof(P2PBlob) * R, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    dev_free(h, d_blob);
    for (int r = 0; r < R && ok; ++r) {
        if (r == me) continue;
        void *pr = nullptr, *ps = nullptr;
        if (cudaIpcOpenMemHandle(&pr, all[r].recv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&ps, all[r].sig, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = 0;
            cudaGetLastError();
            break;
        }
        h->peer_recv[r] = (double4 *)pr;
        h->peer_sig[r] = (unsigned long long *)ps;
    }
    // every rank must take the same path
    CUDA_TRY(h, cudaMemcpyAsync(d_tmp, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(h, g_nccl.AllReduce(d_tmp, d_tmp, 1, ncclInt, ncclMin, h->comm, st));
    int all_ok = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&all_ok, d_tmp, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    if (!all_ok) {
        p2p_teardown(h);
        if (h->halo_mode == 2) return fail(h, PISB_ERR_COMM, "halo_mode 2: peer memory could not be mapped through CUDA IPC");
        return PISB_OK;
    }
    h->p2p_half = half;
    h->p2p_ok = true;
    h->halo_seq = 0;
    return PISB_OK;
}

int halo_exchange_p2p(pisb_t *h) {
    const int R = h->dc.nranks, me = h->dc.rank;
    cudaStream_t st = h->stream;
    const unsigned long long seq = ++h->halo_seq;
    P2PArgs a{};
    a.nranks = R;
    a.me = me;
    a.send_total = h->send_total;
    a.n_ghost = h->n_ghost;
    a.half = h->p2p_half;
    for (int r = 0; r < R; ++r) {
        a.src_off[r] = h->gs_off[r];
        a.cnt[r] = h->gs_cnt[r];
        a.dst_off[r] = h->p2p_dst_off[r];
        a.peer_recv[r] = h->peer_recv[r];
        a.peer_sig[r] = h->peer_sig[r];
    }
    a.my_recv = h->p2p_recv;
    a.my_sig = h->p2p_sig;
    LaunchScope ls(h, PISB_K_HALO);
    if (h->send_total > 0) k_halo_push<<<nblk_capped(h->send_total, TPB, 4), TPB, 0, st>>>(a, seq, h->send_idx.p, h->xt.p);
    k_halo_signal<<<1, 32, 0, st>>>(a, seq, h->flags);
    k_halo_pull<<<nblk_capped(std::max(h->n_ghost, 1), TPB, 4), TPB, 0, st>>>(a, seq, h->ghost_slot.p, h->xt.p, h->xf.p, h->box, h->flags,
                                                                             20000000000LL);
    h->n_launches += 2;
    return check_launch(h, "p2p halo");
}

// Per-step ghost position exchange (no rebuild).  with_flag: the skin-trigger flag of every rank travels in the
// same NCCL group and k_halo_unpack forms the global max (replaces a separate all-reduce).
int halo_exchange(pisb_t *h, bool with_flag) {
    if (h->p2p_ok) return halo_exchange_p2p(h);  // NVLink stores + signals; the flag always rides along
    cudaStream_t st = h->stream;
    const int R = h->dc.nranks, me = h->dc.rank;
    if (h->send_total > 0) {
        LaunchScope ls(h, PISB_K_HALO);
        k_halo_pack<<<nblk(h->send_total, TPB), TPB, 0, st>>>(h->send_total, h->send_idx.p, h->xt.p, h->halo_send.p);
        TRY(check_launch(h, "k_halo_pack"));
    }
    if (with_flag) TRY(dev_reserve(h, h->flag_in, (size_t)R));
    {
        LaunchScope ls(h, PISB_K_HALO);
        const double *sbuf = reinterpret_cast<const double *>(h->halo_send.p);
        double *rbuf = reinterpret_cast<double *>(h->halo_recv.p);
        NCCL_TRY(h, g_nccl.GroupStart());
        for (int r = 0; r < R; ++r) {
            if (r == me) continue;
            if (h->gs_cnt[r] > 0)
                NCCL_TRY(h, g_nccl.Send(sbuf + (size_t)h->gs_off[r] * 4, (size_t)h->gs_cnt[r] * 4, ncclDouble, r, h->comm, st));
            if (h->gr_cnt[r] > 0)
                NCCL_TRY(h, g_nccl.Recv(rbuf + (size_t)h->gr_off[r] * 4, (size_t)h->gr_cnt[r] * 4, ncclDouble, r, h->comm, st));
            if (with_flag) {
                NCCL_TRY(h, g_nccl.Send(h->flags + FLAG_REBUILD, 1, ncclInt, r, h->comm, st));
                NCCL_TRY(h, g_nccl.Recv(h->flag_in.p + r, 1, ncclInt, r, h->comm, st));
            }
        }
        NCCL_TRY(h, g_nccl.GroupEnd());
    }
    if (h->n_ghost > 0 || with_flag) {
        LaunchScope ls(h, PISB_K_HALO);
        k_halo_unpack<<<nblk(std::max(h->n_ghost, 1), TPB), TPB, 0, st>>>(h->n_ghost, h->ghost_slot.p, h->halo_recv.p, h->xt.p,
                                                                         h->xf.p, h->box, with_flag ? h->flag_in.p : nullptr, R, me,
                                                                         h->flags);
        TRY(check_launch(h, "k_halo_unpack"));
    }
    return PISB_OK;
}

int allreduce_thermo(pisb_t *h, size_t nrec) {
    LaunchScope ls(h, PISB_K_REDUCE);
    NCCL_TRY(h, g_nccl.AllReduce(h->thermo_d.p, h->thermo_d.p, nrec * 4, ncclDouble, ncclSum, h->comm, h->stream));
    return PISB_OK;
}

int do_compute_multi(pisb_t *h, int accumulate, double *pe) {
    if (!h->list_valid) TRY(multi_rebuild(h));
    TRY(check_bad_type(h));
    TRY(reserve_thermo(h, 2));
    double *out[3] = {h->f[0].p, h->f[1].p, h->f[2].p};
    const double *acc[3] = {h->f[0].p, h->f[1].p, h->f[2].p};
    TRY(launch_force(h, out, accumulate ? acc : nullptr, h->thermo_d.p));
    TRY(allreduce_thermo(h, 1));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (pe) *pe = h->h_thermo[0].pe;
    h->forces_current = true;
    return PISB_OK;
}

// Spin until k_publish_flags of sequence number `seq` has landed in h_flags (page-locked host memory the kernel writes
// directly).  Every ~64k polls the stream is queried so that a failed launch or a device fault ends the wait.
int wait_published(pisb_t *h, volatile int *pub, int seq) {
    for (unsigned long long it = 1;; ++it) {
        if (__atomic_load_n(pub, __ATOMIC_ACQUIRE) == seq) return PISB_OK;
        if ((it & 0xffffull) == 0) {
            const cudaError_t e = cudaStreamQuery(h->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail(h, PISB_ERR_CUDA, fmt("waiting for the step decision: %s", cudaGetErrorString(e)));
            if (e == cudaSuccess && __atomic_load_n(pub, __ATOMIC_ACQUIRE) != seq)
                return fail(h, PISB_ERR_CUDA, "the stream drained without publishing the step decision");
        }
    }
}

int do_step_nve_multi(pisb_t *h, double dt, int64_t nsteps, pisb_thermo *out) {
    if (!h->list_valid) TRY(multi_rebuild(h));
    TRY(check_bad_type(h));
    const bool trace = std::getenv("PISB_TRACE") != nullptr;
    double t_vv = 0, t_flag = 0, t_reb = 0, t_halo = 0, t_force = 0, t_tail = 0;
    int n_reb = 0;
    const int64_t chunk_max = 4096;
    int64_t done = 0;
    while (done < nsteps) {
        const int64_t m = std::min(chunk_max, nsteps - done);
        TRY(reserve_thermo(h, (size_t)m + 1));
        CUDA_TRY(h, cudaMemsetAsync(h->thermo_d.p, 0, sizeof(pisb_thermo) * (m + 1), h->stream));
        // Bricks the default force kernel serves step with k_force_vv like a single GPU: drift, then per step halo exchange,
        // [rebuild], force + kick + drift of the next step in one launch (owned slots; ghost slots of the other position
        // buffer are filled by the next exchange).
        const bool fused = fused_step_possible(h, true);
        for (int64_t s = 0; s < m; ++s) {
            pisb_thermo *rec = h->thermo_d.p + s;
            double t0 = wall_now();
            if (fused) {
                if (s == 0) TRY(launch_vv(h, false, true, dt, nullptr, nullptr, true));
            } else {
                if (s == 0) TRY(launch_vv(h, false, true, dt, nullptr));
                else TRY(launch_vv(h, true, true, dt, rec - 1));
            }
            double t1 = wall_now();
            // every rank must take the same branch: the skin-trigger flags ride along with the ghost exchange and
            // are max-reduced by k_halo_unpack; the host then reads the decision.  On a rebuild step the ghost
            // exchange is merely redundant (the rebuild re-selects the ghosts).
            TRY(halo_exchange(h, true));
            // The force kernel is launched BEFORE the host knows the decision: it returns at once if the (now global) rebuild
            // flag is set, and the flag travels to the host on a second stream meanwhile -- on the ~80 % of steps without a
            // rebuild the GPU never waits for the host round trip.
            const bool speculate = v2_possible(h) && (h->force_variant == 0 || h->force_variant == 3 || h->force_variant == 6 || h->force_variant == 7);
            double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
            if (speculate) {
                // the decision word the speculative launch and the host read (see FLAG_DECISION) is set, and the flags reach
                // the host, through k_publish_flags: no copy engine, no second stream
                const int seq = ++h->pub_seq;
                k_publish_flags<<<1, 32, 0, h->stream>>>(h->flags, h->h_flags, seq, fused ? 1 : 0);
                h->n_launches += 1;
                if (fused) TRY(launch_force_vv(h, s + 1 < m, dt, rec, h->flags + FLAG_DECISION));
                else TRY(launch_force(h, outp, nullptr, rec, h->flags + FLAG_REBUILD));
                const size_t spec_ev = h->ev_used;  // profiling: event pair of the speculative launch is ev_pool[spec_ev - 1]
                TRY(wait_published(h, h->h_flags + FLAG_COUNT, seq));
                // a launch that turned out to be a no-op is not a force evaluation: book it under the (tiny) reduce class
                if (h->profiling && spec_ev > 0 && h->h_flags[fused ? FLAG_DECISION : FLAG_REBUILD] != 0) h->ev_pool[spec_ev - 1].cls = PISB_K_REDUCE;
            } else {
                CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->flags, sizeof(int) * FLAG_COUNT, cudaMemcpyDeviceToHost, h->stream));
                CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            }
            double t2 = wall_now();
            if (h->h_flags[FLAG_COMM_TIMEOUT]) return fail(h, PISB_ERR_COMM, "timed out waiting for a peer's ghost data (peer-memory halo)");
            const bool reb = h->h_flags[(fused && speculate) ? FLAG_DECISION : FLAG_REBUILD] != 0;
            // a dump frame in flight: its next pieces leave now, while the force kernel runs and no host-visible word is due
            if (!reb && h->dl_pending && h->dl_next < h->dl_queue.size()) TRY(pump_download(h, h->dl_budget));
            if (reb) TRY(multi_rebuild(h));
            double t3 = wall_now();
            if (fused) {
                if (reb || !speculate) {
                    TRY(launch_force_vv(h, s + 1 < m, dt, rec));  // rotates the buffers itself
                } else {  // the speculative launch ran: rotate now
                    for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
                    if (s + 1 < m) swap_position_buffers(h);
                }
            } else {
                if (reb || !speculate) TRY(launch_force(h, outp, nullptr, rec));
                for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
            }
            double t4 = wall_now();
            t_vv += t1 - t0;
            t_flag += t2 - t1;
            (reb ? t_reb : t_halo) += t3 - t2;
            n_reb += reb;
            t_force += t4 - t3;
        }
        double t5 = wall_now();
        if (!fused) TRY(launch_vv(h, true, false, dt, h->thermo_d.p + (m - 1)));
        TRY(allreduce_thermo(h, (size_t)m));
        CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo) * m, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        t_tail += wall_now() - t5;
        if (out) std::memcpy(out + done, h->h_thermo, sizeof(pisb_thermo) * m);
        done += m;
        h->n_steps += m;
    }
    if (trace)
        fprintf(stderr, "[pisb rank %d] %lld steps host wall (ms): vv %.2f flag+sync %.2f rebuild %.2f (%d) halo %.2f force-launch %.2f tail %.2f\n",
                h->dc.rank, (long long)nsteps, 1e3 * t_vv, 1e3 * t_flag, 1e3 * t_reb, n_reb, 1e3 * t_halo, 1e3 * t_force, 1e3 * t_tail);
    h->forces_current = true;
    return PISB_OK;
}

// verlet_step_nvt_nhc (potential.rs:35-58) on bricks.  The thermostat needs the GLOBAL kinetic energy twice per step: the
// chain of every rank is advanced from the same all-reduced numbers by the same one-thread kernel, so the replicas stay
// bit-identical without ever being exchanged.  Per step: chain half step, scaled drift, halo exchange (ghost positions +
// the global rebuild decision), [collective rebuild], force, scaled kick (local KE / virial / PE of the owned atoms), ONE
// all-reduce of the step's 4-number thermo record, chain half step from the reduced KE.
int do_step_nvt_multi(pisb_t *h, double dt, int64_t nsteps, pisb_nhc *chain, int64_t first_step, int64_t total_steps,
                      pisb_thermo *out, double *nhc_energy) {
    if (!chain || chain->chain_size != 3) return fail(h, PISB_ERR_INVALID, "chain must be a 3-link pisb_nhc (pisb_nhc_init)");
    if (nsteps < 0 || total_steps <= 0) return fail(h, PISB_ERR_INVALID, "bad step counts");
    if (nsteps == 0) return PISB_OK;
    if (!h->list_valid) TRY(multi_rebuild(h));
    TRY(check_bad_type(h));
    const int64_t chunk_max = 1024;
    TRY(dev_reserve(h, h->nhc_d, 1));
    TRY(dev_reserve(h, h->nhc_energy_d, (size_t)chunk_max));
    TRY(reserve_thermo(h, (size_t)chunk_max + 2));
    // the state entering the batch: global KE (potential.rs:41) and the global atom count (3N degrees of freedom)
    pisb_thermo *ke0 = h->thermo_d.p + chunk_max;
    {
        LaunchScope ls(h, PISB_K_REDUCE);
        k_observe<<<nblk(h->n, TPB), TPB, 0, h->stream>>>(h->n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p,
                                                          h->f[2].p, h->mass_d.p, h->partials.p, h->ticket, ke0);
        TRY(check_launch(h, "k_observe"));
        const double own = (double)h->n_own;  // rides in the spare record's pe slot through the same all-reduce
        CUDA_TRY(h, cudaMemcpyAsync(&ke0->pe, &own, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        NCCL_TRY(h, g_nccl.AllReduce(ke0, ke0, 4, ncclDouble, ncclSum, h->comm, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, ke0, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    const long long n_global = (long long)std::llround(h->h_thermo[0].pe);
    NhcDev init{};
    init.c = *chain;
    init.scale = 1.0;
    init.ke_last = h->h_thermo[0].ke;
    init.step_index = first_step;
    CUDA_TRY(h, cudaMemcpyAsync(h->nhc_d.p, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    int64_t done = 0;
    while (done < nsteps) {
        const int64_t m = std::min(chunk_max, nsteps - done);
        for (int64_t s = 0; s < m; ++s) {
            pisb_thermo *rec = h->thermo_d.p + s;
            k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, nullptr, n_global, dt, 0, -1, 1, nullptr);
            TRY(launch_vv(h, false, true, dt, nullptr, &h->nhc_d.p->scale));
            TRY(halo_exchange(h, true));
            TRY(read_flags(h));
            if (h->h_flags[FLAG_COMM_TIMEOUT]) return fail(h, PISB_ERR_COMM, "timed out waiting for a peer's ghost data (peer-memory halo)");
            if (h->h_flags[FLAG_REBUILD] != 0) TRY(multi_rebuild(h));
            double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
            TRY(launch_force(h, outp, nullptr, rec));
            for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
            TRY(launch_vv(h, true, false, dt, rec, &h->nhc_d.p->scale));
            {
                LaunchScope ls(h, PISB_K_REDUCE);
                NCCL_TRY(h, g_nccl.AllReduce(rec, rec, 4, ncclDouble, ncclSum, h->comm, h->stream));
            }
            k_nhc_half<<<1, 32, 0, h->stream>>>(h->nhc_d.p, &rec->ke, n_global, dt, 1, -1, (long long)total_steps, h->nhc_energy_d.p + s);
            h->n_launches += 2;
        }
        TRY(check_launch(h, "nvt step (bricks)"));
        CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo) * m, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (out) std::memcpy(out + done, h->h_thermo, sizeof(pisb_thermo) * m);
        if (nhc_energy) CUDA_TRY(h, cudaMemcpy(nhc_energy + done, h->nhc_energy_d.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
        done += m;
        h->n_steps += m;
    }
    NhcDev fin{};
    CUDA_TRY(h, cudaMemcpy(&fin, h->nhc_d.p, sizeof fin, cudaMemcpyDeviceToHost));
    *chain = fin.c;
    h->forces_current = true;
    return PISB_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *pisb_version(void) { return "pisb200 0.1 (sm_100a)"; }

int pisb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *pisb_last_error(pisb_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pisb_create(int device, int n_types, const double *mass, const double *eps, const double *sigma,
                const double *rcut, const unsigned char *present, int shift, double skin, pisb_t **out) {
    if (!out) return fail(nullptr, PISB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_types < 1 || n_types > 64 || !mass || !eps || !sigma || !rcut || !present)
        return fail(nullptr, PISB_ERR_INVALID, "bad potential table arguments");
    if (!(skin >= 0.0)) return fail(nullptr, PISB_ERR_INVALID, "skin must be >= 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, PISB_ERR_NO_DEVICE,
                    fmt("no CUDA device available (%s): libpisb200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, PISB_ERR_INVALID, fmt("device %d out of range (0..%d)", device, ndev - 1));
    pisb_t *h = new pisb_handle();
    h->device = device;
    h->n_types = n_types;
    h->mass.assign(mass, mass + n_types);
    h->eps.assign(eps, eps + n_types * n_types);
    h->sigma.assign(sigma, sigma + n_types * n_types);
    h->rcut.assign(rcut, rcut + n_types * n_types);
    h->present.assign(present, present + n_types * n_types);
    h->shift = shift ? 1 : 0;
    h->skin = skin;
    auto bail = [&](int rc) {
        g_create_error = h->err;
        pisb_destroy(h);
        return rc;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(h, PISB_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_pos, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(h, PISB_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaMalloc((void **)&h->flags, sizeof(int) * FLAG_COUNT) != cudaSuccess ||
        cudaMalloc((void **)&h->ticket, sizeof(unsigned int) * 64) != cudaSuccess ||
        cudaHostAlloc((void **)&h->h_flags, sizeof(int) * (FLAG_COUNT + 2 + 16), cudaHostAllocDefault) != cudaSuccess)
        return bail(fail(h, PISB_ERR_CUDA, "allocating control words failed"));
    cudaMemsetAsync(h->flags, 0, sizeof(int) * FLAG_COUNT, h->stream);
    std::memset(h->h_flags, 0, sizeof(int) * (FLAG_COUNT + 2 + 16));  // incl. the sequence words the host spins on (wait_published)
    h->h_pub = h->h_flags + FLAG_COUNT + 2;
    h->ticket_cap = 64;
    cudaMemsetAsync(h->ticket, 0, sizeof(unsigned int) * 64, h->stream);
    int rc = build_pair_table(h);
    if (rc != PISB_OK) return bail(rc);
    *out = h;
    return PISB_OK;
}

int pisb_destroy(pisb_t *h) {
    if (!h) return PISB_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);  // a download still in flight
    if (h->dl_stream) cudaStreamSynchronize(h->dl_stream);
    free_positions(h, h->xt);
    dev_free(h, h->xf);
    dev_free(h, h->xp);
    dev_free(h, h->tablef_d);
    free_positions(h, h->s_xt);
    dev_free(h, h->xf2);
    dev_free(h, h->dl_pos);
    dev_free(h, h->dl_vel);
    dev_free(h, h->dl_frc);
    dev_free(h, h->dl_ids);
    for (int d = 0; d < 3; ++d) {
        dev_free(h, h->v[d]);
        dev_free(h, h->f[d]);
        dev_free(h, h->g[d]);
        dev_free(h, h->xb[d]);
        dev_free(h, h->s_v[d]);
        dev_free(h, h->s_f[d]);
    }
    dev_free(h, h->id);
    dev_free(h, h->s_id);
    dev_free(h, h->slot_of_id);
    dev_free(h, h->cell_of);
    dev_free(h, h->order);
    dev_free(h, h->nnbr);
    dev_free(h, h->nbr);
    dev_free(h, h->cell_count);
    dev_free(h, h->cell_start);
    dev_free(h, h->tile_sum);
    dev_free(h, h->mass_d);
    dev_free(h, h->partials);
    dev_free(h, h->vel_sums_d);
    dev_free(h, h->st_pos);
    dev_free(h, h->st_vel);
    dev_free(h, h->st_frc);
    dev_free(h, h->st_types);
    dev_free(h, h->table_d);
    dev_free(h, h->nhc_d);
    dev_free(h, h->nhc_energy_d);
    dev_free(h, h->npt_tensors_d);
    if (h->h_npt) cudaFreeHost(h->h_npt);
    if (h->h_energy) cudaFreeHost(h->h_energy);
    dev_free(h, h->thermo_d);
    dev_free(h, h->m_dest);
    dev_free(h, h->m_pig);
    dev_free(h, h->m_cnt);
    dev_free(h, h->m_allcnt);
    dev_free(h, h->m_off);
    dev_free(h, h->send_idx);
    dev_free(h, h->ghost_slot);
    dev_free(h, h->newslot);
    dev_free(h, h->flag_in);
    dev_free(h, h->mig_send);
    dev_free(h, h->mig_recv);
    dev_free(h, h->halo_send);
    dev_free(h, h->halo_recv);
    p2p_teardown(h);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->flags) cudaFree(h->flags);
    if (h->ticket) cudaFree(h->ticket);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->h_thermo) cudaFreeHost(h->h_thermo);
    for (auto &e : h->ev_pool) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    drop_graphs(h);
    if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
    if (h->ev_pos) cudaEventDestroy(h->ev_pos);
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    if (h->up_stream) cudaStreamDestroy(h->up_stream);
    if (h->dl_stream) cudaStreamDestroy(h->dl_stream);
    if (h->ev_dl) cudaEventDestroy(h->ev_dl);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PISB_OK;
}

int pisb_set_box(pisb_t *h, const double *h9, const double *hinv9, const int *pbc3) {
    if (!h) return PISB_ERR_INVALID;
    if (!h9 || !hinv9 || !pbc3) return fail(h, PISB_ERR_INVALID, "NULL box argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    BoxDev b{};
    bool ortho = true;
    for (int k = 0; k < 9; ++k) {
        b.h[k] = h9[k];
        b.hinv[k] = hinv9[k];
        if (!std::isfinite(h9[k]) || !std::isfinite(hinv9[k])) return fail(h, PISB_ERR_INVALID, "box matrix is not finite");
        if (k % 4 != 0 && (h9[k] != 0.0 || hinv9[k] != 0.0)) ortho = false;
    }
    for (int d = 0; d < 3; ++d) b.pbc[d] = pbc3[d] ? 1 : 0;
    b.ortho = ortho ? 1 : 0;
    const bool changed = !h->have_box || std::memcmp(&b, &h->box, sizeof b) != 0;
    h->box = b;
    h->have_box = true;
    if (changed) {
        h->grid_ok = false;
        h->list_valid = false;
        if (h->have_atoms) {
            TRY(set_flag(h, h->xt_flag, 1));  // resident positions were not checked against THIS box: no interior shortcut until a drift wraps them
            TRY(setup_grid(h));
        }
    }
    return PISB_OK;
}

int pisb_upload(pisb_t *h, int64_t n, const double *pos, const double *vel, const double *force,
                const int32_t *types) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    TRY(do_upload(h, n, pos, vel, force, types));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // host buffers may be reused by the caller
    return PISB_OK;
}

int pisb_compute(pisb_t *h, int accumulate, double *pe) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_compute(h, accumulate, pe);
}

int pisb_step_nve(pisb_t *h, double dt, int64_t nsteps, pisb_thermo *out) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_step_nve(h, dt, nsteps, out);
}

int pisb_nhc_init(pisb_nhc *out, double start_temperature, double end_temperature, double tau) {
    if (!out) return PISB_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->chain_size = 3;
    out->start_temperature = start_temperature;
    out->end_temperature = end_temperature;
    out->target_temperature = start_temperature;
    const double q_value = 0.0083144621 * out->target_temperature * (tau * tau);
    double p10 = 1.0;
    for (int i = 0; i < 3; ++i) {
        out->q[i] = q_value / p10;
        p10 *= 10.0;
    }
    return PISB_OK;
}

int pisb_step_nvt_nhc(pisb_t *h, double dt, int64_t nsteps, pisb_nhc *chain, int64_t first_step, int64_t total_steps,
                      pisb_thermo *out, double *nhc_energy) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_step_nvt(h, dt, nsteps, chain, first_step, total_steps, out, nhc_energy);
}

int pisb_mtk_init(pisb_mtk *out, const double *target_pressure9, double tau, int64_t n_atoms, double target_temperature) {
    if (!out || !target_pressure9) return PISB_ERR_INVALID;
    std::memcpy(out->target_pressure, target_pressure9, sizeof out->target_pressure);
    std::memset(out->momentum, 0, sizeof out->momentum);
    out->w = ((double)(3 * n_atoms)) * 0.0083144621 * target_temperature * (tau * tau);  // npt.rs:33
    return PISB_OK;
}

int pisb_step_npt_mtk(pisb_t *h, double dt, int64_t nsteps, pisb_mtk *baro, pisb_nhc *chain, int64_t first_step,
                      int64_t total_steps, pisb_thermo *out, double *ext_energy, double *h9_trace) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_step_npt(h, dt, nsteps, baro, chain, first_step, total_steps, out, ext_energy, h9_trace);
}

int pisb_get_box(pisb_t *h, double *h9, double *hinv9) {
    if (!h) return PISB_ERR_INVALID;
    if (!h->have_box) return fail(h, PISB_ERR_STATE, "get_box before set_box");
    if (h9) std::memcpy(h9, h->box.h, sizeof(double) * 9);
    if (hinv9) std::memcpy(hinv9, h->box.hinv, sizeof(double) * 9);
    return PISB_OK;
}

int pisb_download(pisb_t *h, double *pos, double *vel, double *force) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_download(h, pos, vel, force);
}

int pisb_download_begin(pisb_t *h, double *pos, double *vel, double *force) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_download_begin(h, pos, vel, force);
}

int pisb_download_end(pisb_t *h) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return do_download_end(h);
}

int pisb_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return PISB_ERR_INVALID;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_create_error = fmt("cudaHostRegister: %s", cudaGetErrorString(e));
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PISB_ERR_NO_DEVICE : PISB_ERR_CUDA;
    }
    return PISB_OK;
}

int pisb_host_unregister(void *ptr) {
    if (!ptr) return PISB_ERR_INVALID;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_create_error = fmt("cudaHostUnregister: %s", cudaGetErrorString(e));
        return PISB_ERR_CUDA;
    }
    return PISB_OK;
}

int pisb_thermo_now(pisb_t *h, pisb_thermo *out) {
    if (!h || !out) return PISB_ERR_INVALID;
    if (!h->have_atoms) return fail(h, PISB_ERR_STATE, "thermo before upload");
    CUDA_TRY(h, cudaSetDevice(h->device));
    TRY(reserve_thermo(h, 2));
    {
        LaunchScope ls(h, PISB_K_REDUCE);
        k_observe<<<nblk(h->n, TPB), TPB, 0, h->stream>>>(h->n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p,
                                                          h->f[1].p, h->f[2].p, h->mass_d.p, h->partials.p,
                                                          h->ticket, h->thermo_d.p);
        TRY(check_launch(h, "k_observe"));
    }
    if (h->multi) TRY(allreduce_thermo(h, 1));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *out = h->h_thermo[0];
    return PISB_OK;
}

// velocities.rs:10-59 on the device (pisb_velocity.cuh)
int pisb_start_velocities(pisb_t *h, double temperature, uint64_t seed) {
    if (!h) return PISB_ERR_INVALID;
    if (!h->have_atoms) return fail(h, PISB_ERR_STATE, "start_velocities before upload");
    const double kb = 0.0083144621;  // KB_KJPERMOLEKELVIN, src/constants.rs:3
    for (double m : h->mass) {
        const double sigma = std::sqrt(kb * temperature / m);
        if (!(sigma >= 0.0) || !std::isfinite(sigma))  // rand_distr::Normal::new -> PisError::InvalidDistribution (velocities.rs:25-26)
            return fail(h, PISB_ERR_INVALID, "Invalid normal distribution parameters: standard deviation is invalid");
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    TRY(dev_reserve(h, h->vel_sums_d, 8));
    auto splitmix = [](uint64_t x) {
        x += 0x9E3779B97F4A7C15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    VelInitArgs a{h->n, h->xt.p, h->id.p, h->v[0].p, h->v[1].p, h->v[2].p, h->mass_d.p, kb * temperature,
                  (unsigned long long)splitmix(seed * 0xD1342543DE82EF95ULL + 12345ULL), temperature, h->vel_sums_d.p,
                  h->partials.p, h->ticket};
    const int nb = nblk(h->n, TPB);
    {
        LaunchScope ls(h, PISB_K_INTEGRATE);
        k_vel_create<<<nb, TPB, 0, h->stream>>>(a);
        TRY(check_launch(h, "k_vel_create"));
        if (h->multi) NCCL_TRY(h, g_nccl.AllReduce(a.sums, a.sums, 5, ncclDouble, ncclSum, h->comm, h->stream));
        k_vel_remove_drift<<<nb, TPB, 0, h->stream>>>(a);
        TRY(check_launch(h, "k_vel_remove_drift"));
        if (h->multi) NCCL_TRY(h, g_nccl.AllReduce(a.sums + 5, a.sums + 5, 1, ncclDouble, ncclSum, h->comm, h->stream));
        k_vel_rescale<<<nb, TPB, 0, h->stream>>>(a, kb);
        TRY(check_launch(h, "k_vel_rescale"));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

// The host-buffer step for an atom set the handle already holds (same count, valid list => slot_of_id maps every id to
// its cell-sorted slot): the three arrays travel in chunks of the original atom order and the step is pipelined over PCIe.
//   up_stream   : H2D x, v, F of chunk c                                 (the link stays busy from the first byte)
//   stream      : k_host_load_drift(c) as soon as chunk c is there       (hidden under the copies of chunk c+1)
//   copy_stream : D2H x(t+dt) of chunk c                                 (full duplex: goes down while c+1 comes up)
// then [rebuild], the force pass with the kick (k_force_vv when the default kernel applies), and v, F leave chunk by
// chunk behind k_store_range.  Same kernels' arithmetic, same results as the whole-array sequence below.
static int host_step_pipelined(pisb_t *h, double *pos, double *vel, double *force, double dt, double *pe) {
    h->list_checked = false;  // positions arrive from the host
    const int n = h->n;
    const size_t n3 = (size_t)3 * n;
    TRY(dev_reserve(h, h->st_pos, n3));
    TRY(dev_reserve(h, h->st_vel, n3));
    TRY(dev_reserve(h, h->st_frc, n3));
    TRY(check_bad_type(h));
    TRY(reserve_thermo(h, 2));
    if (!h->up_stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    int chunk = h->host_chunk_atoms > 0 ? h->host_chunk_atoms : std::max(65536, (n + 7) / 8);
    chunk = (chunk + TPB - 1) / TPB * TPB;
    const int nchunk = (n + chunk - 1) / chunk;
    while (h->ev_chunk.size() < (size_t)2 * nchunk + 1) {
        cudaEvent_t e;
        CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_chunk.push_back(e);
    }
    // the staging arrays may still be read by work queued on the main stream (an asynchronous pisb_upload)
    cudaEvent_t ev_start = h->ev_chunk[(size_t)2 * nchunk];
    CUDA_TRY(h, cudaEventRecord(ev_start, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->up_stream, ev_start, 0));
    CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, ev_start, 0));
    const double hs = 0.5 * h->skin;
    const double half_skin2 = h->skin_half2_override >= 0.0 ? h->skin_half2_override : hs * hs;
    const int always = (h->skin > 0.0 && h->skin_half2_override != 0.0) ? 0 : 1;
    for (int c = 0; c < nchunk; ++c) {
        const int o0 = c * chunk, o1 = std::min(n, o0 + chunk);
        const size_t off = (size_t)3 * o0, cnt = (size_t)3 * (o1 - o0);
        cudaEvent_t ev_up = h->ev_chunk[2 * c], ev_dr = h->ev_chunk[2 * c + 1];
        CUDA_TRY(h, cudaMemcpyAsync(h->st_pos.p + off, pos + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(h, cudaMemcpyAsync(h->st_vel.p + off, vel + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(h, cudaMemcpyAsync(h->st_frc.p + off, force + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(h, cudaEventRecord(ev_up, h->up_stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, ev_up, 0));
        {
            LaunchScope ls(h, PISB_K_COPY);
            HostDriftArgs a{o0, o1, h->st_pos.p, h->st_vel.p, h->st_frc.p, h->slot_of_id.p, h->xt.p, h->xf.p,
                            h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->xb[0].p, h->xb[1].p, h->xb[2].p,
                            h->mass_d.p, h->box, dt, dt * dt, half_skin2, always, h->flags, h->xt_flag};
            if (h->box.ortho) k_host_load_drift<true><<<nblk(o1 - o0, TPB), TPB, 0, h->stream>>>(a);
            else k_host_load_drift<false><<<nblk(o1 - o0, TPB), TPB, 0, h->stream>>>(a);
            TRY(check_launch(h, "k_host_load_drift"));
        }
        CUDA_TRY(h, cudaEventRecord(ev_dr, h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, ev_dr, 0));
        CUDA_TRY(h, cudaMemcpyAsync(pos + off, h->st_pos.p + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    h->forces_current = false;
    TRY(launch_rebuild_chain(h));  // runs only if the drift raised the skin trigger; refreshes slot_of_id
    if (fused_step_possible(h)) {
        TRY(launch_force_vv(h, false, dt, h->thermo_d.p));
    } else {
        double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
        TRY(launch_force(h, outp, nullptr, h->thermo_d.p));
        for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
        TRY(launch_vv(h, true, false, dt, h->thermo_d.p));
    }
    for (int c = 0; c < nchunk; ++c) {
        const int o0 = c * chunk, o1 = std::min(n, o0 + chunk);
        const size_t off = (size_t)3 * o0, cnt = (size_t)3 * (o1 - o0);
        cudaEvent_t ev_st = h->ev_chunk[2 * c];
        {
            LaunchScope ls(h, PISB_K_COPY);
            StoreRangeArgs a{o0, o1, h->slot_of_id.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p,
                             h->st_vel.p, h->st_frc.p};
            k_store_range<<<nblk(o1 - o0, TPB), TPB, 0, h->stream>>>(a);
            TRY(check_launch(h, "k_store_range"));
        }
        CUDA_TRY(h, cudaEventRecord(ev_st, h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, ev_st, 0));
        CUDA_TRY(h, cudaMemcpyAsync(vel + off, h->st_vel.p + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->copy_stream));
        CUDA_TRY(h, cudaMemcpyAsync(force + off, h->st_frc.p + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
    TRY(read_flags(h));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
        h->n_builds_host = h->h_flags[FLAG_NBUILDS];
        h->max_nbr = h->h_flags[FLAG_MAXNBR];
    }
    if (h->h_flags[FLAG_MAXNBR] > h->kcap) {
        h->list_valid = false;
        return fail(h, PISB_ERR_CAPACITY, "a neighbour list overflowed during the step: the arrays hold a step computed from a truncated list -- "
                                          "restore x, v, F as they were before the call and repeat it (the capacity grows on the next build)");
    }
    h->n_steps += 1;
    h->forces_current = true;
    if (pe) *pe = h->h_thermo[0].pe;
    return grow_list_if_close(h);
}

int pisb_verlet_step_nve_host(pisb_t *h, int64_t n, double *pos, double *vel, double *force,
                              const int32_t *types, double dt, double *pe) {
    if (!h) return PISB_ERR_INVALID;
    if (!pos || !vel || !force) return fail(h, PISB_ERR_INVALID, "pos/vel/force must be non-NULL");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // types are only re-sent when the atom set is new (they cannot change inside the trait call)
    const bool fresh = !h->have_atoms || h->n != (int)n;
    if (fresh && !types) return fail(h, PISB_ERR_INVALID, "types is NULL on first call");
    if (!fresh && h->host_pipeline && !h->multi && h->have_box && h->grid_ok && h->list_valid && h->slot_of_id.p)
        return host_step_pipelined(h, pos, vel, force, dt, pe);
    TRY(do_upload(h, n, pos, vel, force, fresh ? types : nullptr));
    if (h->multi) return fail(h, PISB_ERR_STATE, "the host-buffer step is a single-GPU entry point");
    // One verlet_step_nve, written out so that x(t+dt) -- final right after the drift -- travels back over PCIe on a
    // second stream while the list check / rebuild, the force kernel and the kick still run.
    TRY(ensure_list(h));
    TRY(check_bad_type(h));
    TRY(reserve_thermo(h, 2));
    const int na = h->n;
    const size_t n3 = (size_t)3 * na;
    TRY(launch_vv(h, false, true, dt, nullptr));
    {
        LaunchScope ls(h, PISB_K_COPY);
        StoreArgs sa{na, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p, h->st_pos.p, nullptr, nullptr};
        k_store_aos<<<nblk(na, TPB), TPB, 0, h->stream>>>(sa);  // on the main stream: ordered before any re-sort
        TRY(check_launch(h, "k_store_aos(pos)"));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_pos, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_pos, 0));
    CUDA_TRY(h, cudaMemcpyAsync(pos, h->st_pos.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->copy_stream));
    TRY(launch_rebuild_chain(h));
    double *outp[3] = {h->g[0].p, h->g[1].p, h->g[2].p};
    TRY(launch_force(h, outp, nullptr, h->thermo_d.p));
    for (int d = 0; d < 3; ++d) std::swap(h->f[d], h->g[d]);
    TRY(launch_vv(h, true, false, dt, h->thermo_d.p));
    {
        LaunchScope ls(h, PISB_K_COPY);
        StoreArgs sa{na, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p, nullptr, h->st_vel.p, h->st_frc.p};
        k_store_aos<<<nblk(na, TPB), TPB, 0, h->stream>>>(sa);
        TRY(check_launch(h, "k_store_aos(vel,frc)"));
    }
    CUDA_TRY(h, cudaMemcpyAsync(vel, h->st_vel.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(force, h->st_frc.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_thermo, h->thermo_d.p, sizeof(pisb_thermo), cudaMemcpyDeviceToHost, h->stream));
    TRY(read_flags(h));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    if (h->h_flags[FLAG_NBUILDS] != h->n_builds_host) {
        h->n_builds_host = h->h_flags[FLAG_NBUILDS];
        h->max_nbr = h->h_flags[FLAG_MAXNBR];
    }
    if (h->h_flags[FLAG_MAXNBR] > h->kcap) {
        h->list_valid = false;
        return fail(h, PISB_ERR_CAPACITY, "a neighbour list overflowed during the step: the arrays hold a step computed from a truncated list -- "
                                          "restore x, v, F as they were before the call and repeat it (the capacity grows on the next build)");
    }
    h->n_steps += 1;
    h->forces_current = true;
    if (pe) *pe = h->h_thermo[0].pe;
    return grow_list_if_close(h);
}

int pisb_neighbours(pisb_t *h, int32_t *nnbr, int32_t *nbr, int64_t cap_per_atom) {
    if (!h || !nnbr) return PISB_ERR_INVALID;
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "neighbours before upload/set_box");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->multi) {
        if (!h->list_valid) TRY(multi_rebuild(h));
    } else {
        TRY(ensure_list(h));
    }
    const int n = h->n;
    std::vector<int> hn(n), hid(n), hl(nbr ? (size_t)h->kcap * h->npad : 0);  // the list itself only when rows are wanted
    CUDA_TRY(h, cudaMemcpyAsync(hn.data(), h->nnbr.p, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(hid.data(), h->id.p, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    if (nbr) CUDA_TRY(h, cudaMemcpyAsync(hl.data(), h->nbr.p, sizeof(int) * hl.size(), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    auto entry = [&](int s, int k) -> int { return hl[nbr_at(k, s, h->npad)]; };
    int64_t total = 0;
    if (h->multi) {
        // rows in owned-slot order (the order pisb_download_owned uses); entries are GLOBAL ids
        std::vector<double4> hx(n);
        CUDA_TRY(h, cudaMemcpy(hx.data(), h->xt.p, sizeof(double4) * n, cudaMemcpyDeviceToHost));
        int o = 0;
        for (int s = 0; s < n; ++s) {
            long long wb;
            std::memcpy(&wb, &hx[s].w, 8);
            if ((wb >> 32) & 1) continue;
            nnbr[o] = hn[s];
            total += hn[s];
            if (nbr) {
                if (hn[s] > cap_per_atom) return fail(h, PISB_ERR_CAPACITY, "cap_per_atom too small");
                for (int k = 0; k < hn[s]; ++k) nbr[(size_t)o * cap_per_atom + k] = hid[entry(s, k)];
            }
            ++o;
        }
        h->total_nbr = total;
        return PISB_OK;
    }
    for (int s = 0; s < n; ++s) {
        const int o = hid[s];
        nnbr[o] = hn[s];
        total += hn[s];
        if (nbr) {
            if (hn[s] > cap_per_atom) return fail(h, PISB_ERR_CAPACITY, fmt("cap_per_atom %lld < list length %d", (long long)cap_per_atom, hn[s]));
            for (int k = 0; k < hn[s]; ++k) nbr[(size_t)o * cap_per_atom + k] = hid[entry(s, k)];
        }
    }
    h->total_nbr = total;
    return PISB_OK;
}

int pisb_list_stats(pisb_t *h, int64_t *out3) {
    if (!h || !out3) return PISB_ERR_INVALID;
    if (!h->have_atoms || !h->have_box) return fail(h, PISB_ERR_STATE, "list_stats before upload/set_box");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->multi) {
        if (!h->list_valid) TRY(multi_rebuild(h));
    } else {
        TRY(ensure_list(h));
    }
    DevBuf<unsigned long long> d;
    TRY(dev_reserve(h, d, 4));
    CUDA_TRY(h, cudaMemsetAsync(d.p, 0, sizeof(unsigned long long) * 4, h->stream));
    ListStatsArgs a{h->n, h->npad, h->xt.p, h->nbr.p, h->nnbr.p, h->box, h->pairs[0], h->table_d.p, h->n_types, d.p};
    if (h->box.ortho) k_list_stats<true><<<nblk(h->n, TPB), TPB, 0, h->stream>>>(a);
    else k_list_stats<false><<<nblk(h->n, TPB), TPB, 0, h->stream>>>(a);
    TRY(check_launch(h, "k_list_stats"));
    unsigned long long host[4] = {0, 0, 0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(host, d.p, sizeof host, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    dev_free(h, d);
    out3[0] = (int64_t)host[0];
    out3[1] = (int64_t)host[1];
    out3[2] = (int64_t)host[0];  // index words stored == listed pairs for per-atom rows
    h->total_nbr = (int64_t)host[0];
    return PISB_OK;
}

int pisb_invalidate_list(pisb_t *h) {
    if (!h) return PISB_ERR_INVALID;
    h->list_valid = false;
    return PISB_OK;
}

int pisb_stats(pisb_t *h, pisb_stats_t *out) {
    if (!h || !out) return PISB_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->n_atoms = h->multi ? h->n_own : h->n;
    out->n_ghost = h->multi ? h->n_ghost : 0;
    for (int d = 0; d < 3; ++d) out->n_cells[d] = h->grid.n[d];
    out->list_capacity = h->kcap;
    out->max_neighbours = h->max_nbr;
    out->capacity_growths = h->n_capacity_growths;
    out->n_builds = h->n_builds_host;
    out->n_steps = h->n_steps;
    out->n_launches = h->n_launches;
    out->device_bytes = h->device_bytes;
    out->missing_type_pairs = h->missing_type_pairs;
    return PISB_OK;
}

int pisb_set_profiling(pisb_t *h, int enable) {
    if (!h) return PISB_ERR_INVALID;
    drain_events(h);
    h->profiling = enable != 0;
    return PISB_OK;
}

int pisb_timings(pisb_t *h, double *ms, int64_t *launches) {
    if (!h) return PISB_ERR_INVALID;
    cudaSetDevice(h->device);
    drain_events(h);
    for (int k = 0; k < PISB_K_COUNT; ++k) {
        if (ms) ms[k] = h->t_ms[k];
        if (launches) launches[k] = h->t_n[k];
    }
    return PISB_OK;
}

int pisb_timings_reset(pisb_t *h) {
    if (!h) return PISB_ERR_INVALID;
    drain_events(h);
    for (int k = 0; k < PISB_K_COUNT; ++k) h->t_ms[k] = 0.0, h->t_n[k] = 0;
    return PISB_OK;
}


// ---- multi-GPU entry points -------------------------------------------------------------------
int pisb_comm_unique_id(void *out, int nbytes) {
    if (!out || nbytes < (int)sizeof(ncclUniqueId)) return fail(nullptr, PISB_ERR_INVALID, "unique-id buffer must hold 128 bytes");
    TRY(load_nccl(nullptr));
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, PISB_ERR_COMM, fmt("ncclGetUniqueId: %s", g_nccl.GetErrorString(r)));
    std::memcpy(out, &id, sizeof id);
    return PISB_OK;
}

int pisb_comm_init(pisb_t *h, int rank, int nranks, const void *unique_id, const int *grid3) {
    if (!h || !unique_id || !grid3) return PISB_ERR_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, PISB_ERR_INVALID, "bad rank / nranks");
    for (int d = 0; d < 3; ++d)
        if (grid3[d] != 1 && grid3[d] != 2) return fail(h, PISB_ERR_INVALID, "bricks per dimension must be 1 or 2 (one NVSwitch node)");
    if (grid3[0] * grid3[1] * grid3[2] != nranks) return fail(h, PISB_ERR_INVALID, "grid does not match nranks");
    CUDA_TRY(h, cudaSetDevice(h->device));
    TRY(load_nccl(h));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    NCCL_TRY(h, g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->multi = true;
    h->dc = Decomp{};
    for (int d = 0; d < 3; ++d) h->dc.P[d] = grid3[d];
    h->dc.rank = rank;
    h->dc.nranks = nranks;
    h->dc.b[0] = rank % grid3[0];
    h->dc.b[1] = (rank / grid3[0]) % grid3[1];
    h->dc.b[2] = rank / (grid3[0] * grid3[1]);
    CUDA_TRY(h, cudaHostAlloc((void **)&h->h_counts, sizeof(int) * (nranks * nranks + 2), cudaHostAllocDefault));  // + the sequence word of k_publish_words
    std::memset(h->h_counts, 0, sizeof(int) * (nranks * nranks + 2));
    h->grid_ok = false;
    h->list_valid = false;
    return PISB_OK;
}

int pisb_upload_owned(pisb_t *h, int64_t n_own, const double *pos, const double *vel, const double *force,
                      const int32_t *types, const int32_t *gids) {
    if (!h) return PISB_ERR_INVALID;
    if (!h->multi) return fail(h, PISB_ERR_STATE, "pisb_comm_init must precede pisb_upload_owned");
    if (!h->have_box) return fail(h, PISB_ERR_STATE, "pisb_set_box must precede pisb_upload_owned");
    if (n_own < 0 || n_own > 2000000000LL || (n_own > 0 && (!pos || !types || !gids)))
        return fail(h, PISB_ERR_INVALID, "bad pisb_upload_owned arguments");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int n = (int)n_own;
    // capacity: owned + ghost shell, with head-room for density fluctuations and migration
    double ratio = 1.0;
    const double gw = (h->max_rcut + h->skin) * 1.05 + h->skin;
    for (int d = 0; d < 3; ++d)
        if (h->dc.P[d] > 1) {
            const double w = h->box.h[4 * d] / h->dc.P[d];
            ratio *= (w + 2.0 * gw) / w;
        }
    const int cap = (int)std::min(2.0e9, (2.0 * ratio - 1.0) * 1.25 * std::max(n, 1024) + 4096.0);  // owned + stale + fresh ghosts
    if (cap > h->ncap_atoms) {
        h->ncap_atoms = cap;
        h->kcap = 0;
        if (h->p2p_tried) {  // receive buffers are sized by the capacity: map them again at the next rebuild
            p2p_teardown(h);
            h->p2p_tried = false;
        }
    }
    h->n = n;
    h->n_own = n;
    h->n_ghost = 0;
    h->have_atoms = true;
    h->list_valid = false;
    h->grid_ok = false;
    TRY(reserve_atoms(h, h->ncap_atoms));
    const size_t n3 = (size_t)3 * std::max(n, 1);
    TRY(dev_reserve(h, h->st_pos, n3));
    TRY(dev_reserve(h, h->st_vel, n3));
    TRY(dev_reserve(h, h->st_frc, n3));
    TRY(dev_reserve(h, h->st_types, (size_t)std::max(n, 1)));
    TRY(dev_reserve(h, h->m_dest, (size_t)std::max(n, 1)));  // staging for gids
    if (n > 0) {
        CUDA_TRY(h, cudaMemcpyAsync(h->st_pos.p, pos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        if (vel) CUDA_TRY(h, cudaMemcpyAsync(h->st_vel.p, vel, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        if (force) CUDA_TRY(h, cudaMemcpyAsync(h->st_frc.p, force, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(h->st_types.p, types, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    count_missing_pairs(h, n_own, types);  // this rank's atoms
        CUDA_TRY(h, cudaMemcpyAsync(h->m_dest.p, gids, sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
        LaunchScope ls(h, PISB_K_COPY);
        LoadArgs la{n, h->st_pos.p, vel ? h->st_vel.p : nullptr, force ? h->st_frc.p : nullptr, h->st_types.p, nullptr,
                    h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p, h->f[1].p, h->f[2].p, h->id.p, h->n_types,
                    h->flags, nullptr, h->box, h->m_dest.p, 1, h->xt_flag};
        TRY(set_flag(h, h->xt_flag, 0));
        k_load_aos<<<nblk(n, TPB), TPB, 0, h->stream>>>(la);
        TRY(check_launch(h, "k_load_aos"));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->forces_current = false;
    return PISB_OK;
}

int pisb_download_owned(pisb_t *h, int64_t cap, double *pos, double *vel, double *force, int32_t *gids,
                        int64_t *n_out) {
    if (!h || !n_out || !gids) return PISB_ERR_INVALID;
    if (!h->multi || !h->have_atoms) return fail(h, PISB_ERR_STATE, "pisb_download_owned needs multi-GPU mode and uploaded atoms");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int n = h->n, no = h->n_own;
    if (no > cap) return fail(h, PISB_ERR_CAPACITY, "pisb_download_owned: cap too small");
    const size_t n3 = (size_t)3 * std::max(no, 1);
    if (pos) TRY(dev_reserve_grow(h, h->st_pos, n3));
    if (vel) TRY(dev_reserve_grow(h, h->st_vel, n3));
    if (force) TRY(dev_reserve_grow(h, h->st_frc, n3));
    TRY(dev_reserve_grow(h, h->st_types, (size_t)std::max(no, 1)));  // staging for the ids
    TRY(dev_reserve(h, h->m_cnt, (size_t)std::max(h->dc.nranks, 1)));
    CUDA_TRY(h, cudaMemsetAsync(h->m_cnt.p, 0, sizeof(int), h->stream));
    {
        LaunchScope ls(h, PISB_K_COPY);
        k_store_owned<<<nblk(std::max(n, 1), TPB), TPB, 0, h->stream>>>(n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p,
                                                                       h->f[1].p, h->f[2].p, h->id.p, h->m_cnt.p,
                                                                       pos ? h->st_pos.p : nullptr, vel ? h->st_vel.p : nullptr,
                                                                       force ? h->st_frc.p : nullptr, h->st_types.p);
        TRY(check_launch(h, "k_store_owned"));
    }
    if (pos) CUDA_TRY(h, cudaMemcpyAsync(pos, h->st_pos.p, sizeof(double) * 3 * no, cudaMemcpyDeviceToHost, h->stream));
    if (vel) CUDA_TRY(h, cudaMemcpyAsync(vel, h->st_vel.p, sizeof(double) * 3 * no, cudaMemcpyDeviceToHost, h->stream));
    if (force) CUDA_TRY(h, cudaMemcpyAsync(force, h->st_frc.p, sizeof(double) * 3 * no, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(gids, h->st_types.p, sizeof(int) * no, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *n_out = no;
    return PISB_OK;
}

// pisb_download_owned_begin: the same rows, asynchronously.  The owned atoms are compacted into snapshot buffers on the main
// stream (a few hundred microseconds), the copies run on a stream of their own behind an event, and the next batch may start
// at once; pisb_download_end waits for the frame.  n_out is known on the host (the owned count changes only in a rebuild).
int pisb_download_owned_begin(pisb_t *h, int64_t cap, double *pos, double *vel, double *force, int32_t *gids, int64_t *n_out) {
    if (!h || !n_out || !gids) return PISB_ERR_INVALID;
    if (!h->multi || !h->have_atoms) return fail(h, PISB_ERR_STATE, "pisb_download_owned_begin needs multi-GPU mode and uploaded atoms");
    if (h->dl_pending) return fail(h, PISB_ERR_STATE, "a download is already in flight: call pisb_download_end first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int n = h->n, no = h->n_own;
    if (no > cap) return fail(h, PISB_ERR_CAPACITY, "pisb_download_owned_begin: cap too small");
    if (!h->dl_stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->dl_stream, cudaStreamNonBlocking));
    if (!h->ev_dl) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_dl, cudaEventDisableTiming));
    const size_t n3 = (size_t)3 * std::max(no, 1);
    if (pos) TRY(dev_reserve_grow(h, h->dl_pos, n3));
    if (vel) TRY(dev_reserve_grow(h, h->dl_vel, n3));
    if (force) TRY(dev_reserve_grow(h, h->dl_frc, n3));
    TRY(dev_reserve_grow(h, h->dl_ids, (size_t)std::max(no, 1)));
    TRY(dev_reserve(h, h->m_cnt, (size_t)std::max(h->dc.nranks, 1)));
    CUDA_TRY(h, cudaMemsetAsync(h->m_cnt.p, 0, sizeof(int), h->stream));
    {
        LaunchScope ls(h, PISB_K_COPY);
        k_store_owned<<<nblk(std::max(n, 1), TPB), TPB, 0, h->stream>>>(n, h->xt.p, h->v[0].p, h->v[1].p, h->v[2].p, h->f[0].p,
                                                                       h->f[1].p, h->f[2].p, h->id.p, h->m_cnt.p,
                                                                       pos ? h->dl_pos.p : nullptr, vel ? h->dl_vel.p : nullptr,
                                                                       force ? h->dl_frc.p : nullptr, h->dl_ids.p);
        TRY(check_launch(h, "k_store_owned"));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_dl, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->dl_stream, h->ev_dl, 0));
    h->dl_pending = true;
    // The frame is queued in pieces of <= 2 MB; the brick step loop sends a budget of them after every step's decision has
    // reached the host (total / dl_spread_steps per step), pisb_download_end sends what is left.  In one piece the frame would
    // sit in front of every small device-to-host word of the next steps: at 8 GPUs (eight frames into one host, ~12 GB/s each)
    // that made the three host reads of a rebuild 0.8 ms each instead of 0.07.
    h->dl_queue.clear();
    h->dl_next = 0;
    size_t total = 0;
    auto enqueue = [&](void *dst, const void *src, size_t bytes) {
        const size_t piece = (size_t)2 << 20;
        for (size_t off = 0; off < bytes; off += piece)
            h->dl_queue.push_back({(char *)dst + off, (const char *)src + off, std::min(piece, bytes - off)});
        total += bytes;
    };
    enqueue(gids, h->dl_ids.p, sizeof(int) * (size_t)no);
    if (pos) enqueue(pos, h->dl_pos.p, sizeof(double) * 3 * (size_t)no);
    if (vel) enqueue(vel, h->dl_vel.p, sizeof(double) * 3 * (size_t)no);
    if (force) enqueue(force, h->dl_frc.p, sizeof(double) * 3 * (size_t)no);
    h->dl_budget = total / (size_t)std::max(h->dl_spread_steps, 1) + 1;
    *n_out = no;
    return PISB_OK;
}

// Test hook (multi-GPU): global ids of the owned atoms in device slot order = the row order of pisb_neighbours.
int pisb_owned_ids(pisb_t *h, int64_t cap, int32_t *gids, int64_t *n_out) {
    if (!h || !n_out || !gids) return PISB_ERR_INVALID;
    if (!h->multi || !h->have_atoms) return fail(h, PISB_ERR_STATE, "pisb_owned_ids needs multi-GPU mode and uploaded atoms");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->list_valid) TRY(multi_rebuild(h));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const int n = h->n;
    std::vector<double4> hx(n);
    std::vector<int> hid(n);
    CUDA_TRY(h, cudaMemcpy(hx.data(), h->xt.p, sizeof(double4) * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(hid.data(), h->id.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    int64_t o = 0;
    for (int s = 0; s < n; ++s) {
        long long wb;
        std::memcpy(&wb, &hx[s].w, 8);
        if ((wb >> 32) & 1) continue;
        if (o >= cap) return fail(h, PISB_ERR_CAPACITY, "pisb_owned_ids: cap too small");
        gids[o++] = hid[s];
    }
    *n_out = o;
    return PISB_OK;
}

void *pisb_stream(pisb_t *h) { return h ? (void *)h->stream : nullptr; }

int pisb_synchronize(pisb_t *h) {
    if (!h) return PISB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return PISB_OK;
}

int pisb_set_option(pisb_t *h, const char *name, double value) {
    if (!h || !name) return PISB_ERR_INVALID;
    if (!std::strcmp(name, "list_capacity")) {
        h->kcap_user = value > 0 ? (int)value : 0;
        h->kcap = 0;
        h->list_valid = false;
        return PISB_OK;
    }
    if (!std::strcmp(name, "cuda_graphs")) {
        h->use_graphs = value != 0.0 ? 1 : 0;
        return PISB_OK;
    }
    if (!std::strcmp(name, "dump_spread_steps")) {
        h->dl_spread_steps = value >= 1.0 ? (int)value : 1;
        return PISB_OK;
    }
    if (!std::strcmp(name, "halo_mode")) {
        h->halo_mode = (int)value;
        return PISB_OK;
    }
    if (!std::strcmp(name, "host_pipeline")) {
        h->host_pipeline = value != 0.0 ? 1 : 0;
        return PISB_OK;
    }
    if (!std::strcmp(name, "host_chunk_atoms")) {
        h->host_chunk_atoms = value > 0 ? (int)value : 0;
        return PISB_OK;
    }
    if (!std::strcmp(name, "fuse_vv")) {
        h->fuse_vv = value != 0.0 ? 1 : 0;
        return PISB_OK;
    }
    if (!std::strcmp(name, "force_variant")) {
        const int v = (int)value;
        if (v < 0 || v == 4 || v == 5 || v > 7) return fail(h, PISB_ERR_INVALID, fmt("force_variant %d does not exist (0-3, 6, 7)", v));
        h->force_variant = v;
        return PISB_OK;
    }
    if (!std::strcmp(name, "build_variant") || !std::strcmp(name, "cell_div")) {
        if (name[0] == 'b') h->build_variant = (int)value;
        else h->cell_div = (int)value;
        h->list_valid = false;
        h->grid_ok = false;
        return PISB_OK;
    }
    return fail(h, PISB_ERR_INVALID, fmt("unknown option '%s'", name));
}

}  // extern "C"
