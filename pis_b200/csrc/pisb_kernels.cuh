// pisb_kernels.cuh -- the sm_100a kernels of the hot path (single translation unit with pisb_sim.cu).
//
// Device data layout (all arrays in CELL-SORTED slot order; `id[slot]` = original atom index):
//   xt  : double4 {x, y, z, type-bits}   one 32-byte sector per atom: a neighbour gather costs
//                                        exactly one sector and carries the type for the pair table;
//                                        index -1 holds a NaN record (what a list pad gathers)
//   v*, f*, g* : SoA doubles             streamed, never gathered (f = force at current x,
//                                        g = force at previous x, swapped by the host every step)
//   xb* : SoA doubles                    positions at the last list build (skin trigger)
//   xf  : float4                          FP32 shadow of the wrapped position (+ type / ghost bits) for the pre-filter
//   nbr : int32 [K_cap/4][n_pad][4]      FULL Verlet list, transposed in K-tiles of 4 (see nbr_at): the neighbours
//                                        k..k+3 of atom i are one aligned int4, a warp reads 512 contiguous bytes;
//                                        the last tile of a row is closed with pads (sign bit set)
//   cell_start : int32 [n_cells+2]       CSR of the cell-sorted order (x fastest, like cell_index)
//
// Kernel variants (pisb_set_option "force_variant" / "build_variant"): the same in/out decision for every pair in all of
// them; v1 / v2 keep the reference's operation order per pair (bit-identical pair terms), the lean kernels shorten it.
//   k_force (v1)      plain all-FP64 loop, round(); also the only path for triclinic / non-periodic boxes
//   k_force_v2        FP32 pre-filter + shared-memory compaction queue + exact FP64 pair terms
//   k_force_v3        all-FP64 LEAN loop (short-way distance + FP64 guard band, Newton reciprocal, factored constants),
//                     int4 index tiles prefetched, 4 x 256-bit gathers in flight, interior-warp shortcut, 64-register bound
//   k_force_vv        DEFAULT above 75k atoms: k_force_v3's loop + the velocity-Verlet kick and drift in the epilogue
//   k_force_q         DEFAULT up to 75k atoms: four lanes per atom, lane l takes entry l of every K-tile, integrator in the epilogue
//   k_force_split<S>  S lanes per atom taking whole K-tiles (round-1 small-system kernel, force_variant 6)
//   k_build_list (v1) all-FP64 27-cell scan;  k_build_list_v2  scalar FP32 pre-filter over contiguous x-rows;
//   k_build_list_v3   DEFAULT: packed-FP32 pair records, bit-mask append, per-pair bands for multi-type tables
#pragma once
#include "pisb_device.cuh"

namespace pisb {

constexpr int TPB = 256;        // streaming kernels
constexpr int TPB_FORCE = 128;  // force / build kernels
constexpr int TPB_VV = 512;     // k_force_v3 / k_force_vv: 2 resident blocks per SM at the 64-register bound (1.130 against 1.149 ms
                                // per launch with 128; 256: 1.145, 1024: 1.180 -- profiles/r02_ab_force_block_sched_4M_43K.jsonl)

// flags[] (device ints)
// FLAG_DECISION: multi-GPU fused steps -- a stream-ordered copy of FLAG_REBUILD taken after the halo exchange.  The
// speculative k_force_vv launch and the host both read THIS word: the kernel's own drift raises FLAG_REBUILD for the next
// step while the launch is still running.
// FLAG_UNWRAPPED_A / _B: one word per position buffer (xt / s_xt trade places in fused steps): set by an upload that found a
// coordinate outside [0, L) in that buffer, cleared by the drift that writes wrapped positions into it.  While it is set the
// force kernels take the minimum-image path for EVERY warp: the interior shortcut uses raw differences and is only valid
// when every position (the atom's and its neighbours') is the wrapped one.
enum { FLAG_REBUILD = 0, FLAG_MAXNBR = 1, FLAG_NBUILDS = 2, FLAG_BADTYPE = 3, FLAG_COMM_TIMEOUT = 4, FLAG_DECISION = 5,
       FLAG_UNWRAPPED_A = 6, FLAG_UNWRAPPED_B = 7, FLAG_COUNT = 8 };

// Neighbour list layout: K-tiles of 4.  Entry (k, i) lives at ((k/4)*npad + i)*4 + k%4, so the four
// neighbours k..k+3 of atom i are one aligned int4 and a warp reads 512 contiguous bytes per tile.
__host__ __device__ __forceinline__ size_t nbr_at(int k, int i, int npad) {
    return ((size_t)(k >> 2) * (size_t)npad + (size_t)i) * 4 + (size_t)(k & 3);
}

// Pads: an entry with the sign bit set is no neighbour.  Every build closes the last K-tile of a row with them, and the
// position buffers carry a NaN record at index -1 (reserve_positions), so the hot loop reads whole int4 tiles without a
// `k < nn` test: one IMNMX turns a pad into a gather of that record, whose r2 is NaN and fails every range test.
constexpr int NBR_PAD_BIT = (int)0x80000000u;
__device__ __forceinline__ void nbr_close_row(int *nbr, int i, int npad, int cnt, int kcap) {
    for (int k = cnt; (k & 3) && k < kcap; ++k) nbr[nbr_at(k, i, npad)] = NBR_PAD_BIT;
}

struct Grid {
    int n[3];      // cells per dimension
    int lo[3];     // stencil offsets lo..hi per dim (dedupes n<3)
    int hi[3];
    int ncell;
    // multi-GPU: dims that are decomposed over ranks are binned brick-locally and non-periodically
    // (ghost atoms supply the periodic / neighbouring images); positions stay GLOBAL wrapped coordinates.
    int local[3];
    double center[3];    // brick centre (global coordinate)
    double half[3];      // half extent of the extended brick: brick/2 + ghost width
    double inv_edge[3];  // cells per unit length
};

// ghost flag lives above the type in the w lane of the position record (and bit 30 of xf.w)
__device__ __forceinline__ bool is_ghost(double w) { return (__double_as_longlong(w) >> 32) & 1; }
__device__ __forceinline__ double type_ghost_as_double(int t, bool ghost) {
    return __longlong_as_double((long long)t | ((long long)(ghost ? 1 : 0) << 32));
}
// "dead" slots (atoms that migrated away, stale ghosts) carry bit 33 (and the ghost bit, so every kernel
// skips them); k_bin files them in a sentinel bucket behind the last cell and the sort drops them.
__device__ __forceinline__ bool is_dead(double w) { return (__double_as_longlong(w) >> 33) & 1; }
__device__ __forceinline__ double mark_dead(double w) { return __longlong_as_double(__double_as_longlong(w) | (3LL << 32)); }
constexpr int XF_GHOST_BIT = 1 << 30;
__device__ __forceinline__ bool xf_is_ghost(const float4 &x) { return (__float_as_int(x.w) & XF_GHOST_BIT) != 0; }
__device__ __forceinline__ int xf_type(const float4 &x) { return __float_as_int(x.w) & (XF_GHOST_BIT - 1); }

// Cell coordinates of a position (shared by k_bin and the build kernels so they always agree).
template <bool ORTHO>
__device__ __forceinline__ void cell_coords(const BoxDev &box, const Grid &g, double x, double y, double z, int *c) {
    double s[3];
    matvec<ORTHO>(box.hinv, x, y, z, s[0], s[1], s[2]);
    const double r[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (g.local[d]) {
            double dc = r[d] - g.center[d];
            dc -= box.h[4 * d] * rint(dc * box.hinv[4 * d]);
            int cd = (int)floor((dc + g.half[d]) * g.inv_edge[d]);
            c[d] = min(max(cd, 0), g.n[d] - 1);
        } else {
            double sd = s[d];
            if (box.pbc[d]) sd = sd - floor(sd);
            // saturating conversion like Rust `as usize` (NaN/negative -> 0)
            unsigned long long cd = __double2ull_rz(floor(sd * (double)g.n[d]));
            if (cd >= (unsigned long long)g.n[d])
                cd = box.pbc[d] ? cd % (unsigned long long)g.n[d] : (unsigned long long)(g.n[d] - 1);
            c[d] = (int)cd;
        }
    }
}

// FP32 shadow of a position for the v2 pre-filter: the WRAPPED coordinate (the filter only needs an
// approximation consistent with the minimum image) + the type bits in w.
__device__ __forceinline__ float4 make_xf(const BoxDev &b, const double4 &x) {
    double c[3] = {x.x, x.y, x.z};
    if (b.ortho) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (b.pbc[d]) {
                double s = c[d] * b.hinv[4 * d];
                s -= floor(s);
                c[d] = b.h[4 * d] * s;
            }
        }
    }
    return make_float4((float)c[0], (float)c[1], (float)c[2],
                       __int_as_float(type_of(x.w) | (is_ghost(x.w) ? XF_GHOST_BIT : 0)));
}

// ------------------------------------------------------------------------------------------------
// Host AoS <-> device layout.  Replaces nothing in the reference: it is the boundary copy.
// `order_by_id` : slot_of_id (keep the current cell-sorted order) or nullptr (identity order).
// ------------------------------------------------------------------------------------------------
struct LoadArgs {
    int n;
    const double *pos, *vel, *frc;  // AoS staging (device), vel/frc may be null
    const int *types;               // 1-based types in ORIGINAL order (device copy kept by the handle)
    const int *slot_of_id;          // may be null => identity
    double4 *xt;
    double *vx, *vy, *vz, *fx, *fy, *fz;
    int *id;
    int n_types;
    int *flags;
    float4 *xf;   // written when the current order/list is kept (null otherwise: the rebuild writes it)
    BoxDev box;
    const int *gids;  // multi-GPU: global ids of the uploaded (owned) atoms; null => id = upload index
    int have_box;     // box is valid: positions can be tested against / wrapped into it
    int unwrapped_flag;  // flags[] word of the position buffer being written
};

__global__ void __launch_bounds__(TPB) k_load_aos(LoadArgs a) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;  // original index
    if (o >= a.n) return;
    int s = a.slot_of_id ? a.slot_of_id[o] : o;
    double4 x;
    x.x = a.pos[3 * (size_t)o];
    x.y = a.pos[3 * (size_t)o + 1];
    x.z = a.pos[3 * (size_t)o + 2];
    int t = a.types[o];
    if (t < 1 || t > a.n_types) {
        a.flags[FLAG_BADTYPE] = o + 1;
        t = 1;
    }
    x.w = type_as_double(t);
    if (!a.have_box) {
        if (o == 0) a.flags[a.unwrapped_flag] = 1;  // no box yet: nothing is known about these coordinates
    } else if (a.box.ortho) {
        if (a.gids) {
            // multi-GPU: positions are GLOBAL WRAPPED coordinates everywhere (ownership, ghost selection, interior test)
            wrap_pos<true>(a.box, x.x, x.y, x.z);
        } else {
            // single GPU: pisb_compute must not move the caller's positions (compute_potential leaves them alone and
            // Sum r.F uses them as given); record that this buffer holds unwrapped coordinates instead
            const bool out = (a.box.pbc[0] && !(x.x >= 0.0 && x.x < a.box.h[0])) || (a.box.pbc[1] && !(x.y >= 0.0 && x.y < a.box.h[4])) ||
                             (a.box.pbc[2] && !(x.z >= 0.0 && x.z < a.box.h[8]));
            if (out) a.flags[a.unwrapped_flag] = 1;
        }
    }
    a.xt[s] = x;
    if (a.xf) a.xf[s] = make_xf(a.box, x);
    a.vx[s] = a.vel ? a.vel[3 * (size_t)o] : 0.0;
    a.vy[s] = a.vel ? a.vel[3 * (size_t)o + 1] : 0.0;
    a.vz[s] = a.vel ? a.vel[3 * (size_t)o + 2] : 0.0;
    a.fx[s] = a.frc ? a.frc[3 * (size_t)o] : 0.0;
    a.fy[s] = a.frc ? a.frc[3 * (size_t)o + 1] : 0.0;
    a.fz[s] = a.frc ? a.frc[3 * (size_t)o + 2] : 0.0;
    a.id[s] = a.gids ? a.gids[o] : o;
}

struct StoreArgs {
    int n;
    const double4 *xt;
    const double *vx, *vy, *vz, *fx, *fy, *fz;
    const int *id;
    double *pos, *vel, *frc;  // AoS staging (device), any may be null
};

__global__ void __launch_bounds__(TPB) k_store_aos(StoreArgs a) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    size_t o = (size_t)a.id[s];
    if (a.pos) {
        double4 x = a.xt[s];
        a.pos[3 * o] = x.x;
        a.pos[3 * o + 1] = x.y;
        a.pos[3 * o + 2] = x.z;
    }
    if (a.vel) {
        a.vel[3 * o] = a.vx[s];
        a.vel[3 * o + 1] = a.vy[s];
        a.vel[3 * o + 2] = a.vz[s];
    }
    if (a.frc) {
        a.frc[3 * o] = a.fx[s];
        a.frc[3 * o + 1] = a.fy[s];
        a.frc[3 * o + 2] = a.fz[s];
    }
}

// ------------------------------------------------------------------------------------------------
// Chunked boundary kernels of the host-buffer step (pisb_verlet_step_nve_host): the trait call moves 288 MB each
// way per step at 4M atoms, so the step is cut into chunks of the ORIGINAL atom order and pipelined over PCIe --
// chunk c is drifted as soon as its x, v, F have arrived and its x(t+dt) travels back while chunk c+1 is still
// on its way up (full duplex); after the force pass the velocities and forces leave chunk by chunk the same way.
//   k_host_load_drift : k_load_aos (kept slot order) + k_vv<drift> for the ids [o0, o1); x(t+dt) also goes back
//                       into the staging array, in place.  Arithmetic = k_vv's, operation for operation.
//   k_store_range     : k_store_aos for the ids [o0, o1) (gather through slot_of_id, contiguous staging writes)
// ------------------------------------------------------------------------------------------------
struct HostDriftArgs {
    int o0, o1;
    double *pos;                 // AoS staging (device): x(t) in, x(t+dt) out
    const double *vel, *frc;     // AoS staging (device)
    const int *slot_of_id;
    double4 *xt;
    float4 *xf;
    double *vx, *vy, *vz, *fx, *fy, *fz;
    const double *xbx, *xby, *xbz;
    const double *mass;
    BoxDev box;
    double dt, dt2, half_skin2;
    int always_rebuild;
    int *flags;
    int unwrapped_flag;  // word of the position buffer written (cleared: the drift wraps)
};

template <bool ORTHO>
__global__ void __launch_bounds__(TPB) k_host_load_drift(HostDriftArgs a) {
    const int o = a.o0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= a.o1) return;
    const int s = a.slot_of_id[o];
    const size_t o3 = 3 * (size_t)o;
    double4 x;
    x.x = a.pos[o3];
    x.y = a.pos[o3 + 1];
    x.z = a.pos[o3 + 2];
    x.w = a.xt[s].w;  // the type does not change inside the trait call
    const double vx = a.vel[o3], vy = a.vel[o3 + 1], vz = a.vel[o3 + 2];
    const double fx = a.frc[o3], fy = a.frc[o3 + 1], fz = a.frc[o3 + 2];
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (!a.always_rebuild) {
        bx = a.xbx[s];
        by = a.xby[s];
        bz = a.xbz[s];
    }
    const double m = a.mass[type_of(x.w) - 1];
    const double ax = __ddiv_rn(fx, m), ay = __ddiv_rn(fy, m), az = __ddiv_rn(fz, m);
    x.x = __dadd_rn(x.x, __dadd_rn(__dmul_rn(vx, a.dt), __dmul_rn(__dmul_rn(ax, 0.5), a.dt2)));
    x.y = __dadd_rn(x.y, __dadd_rn(__dmul_rn(vy, a.dt), __dmul_rn(__dmul_rn(ay, 0.5), a.dt2)));
    x.z = __dadd_rn(x.z, __dadd_rn(__dmul_rn(vz, a.dt), __dmul_rn(__dmul_rn(az, 0.5), a.dt2)));
    wrap_pos<ORTHO>(a.box, x.x, x.y, x.z);
    a.xt[s] = x;
    a.xf[s] = make_float4((float)x.x, (float)x.y, (float)x.z, __int_as_float(type_of(x.w)));
    a.vx[s] = vx;
    a.vy[s] = vy;
    a.vz[s] = vz;
    a.fx[s] = fx;
    a.fy[s] = fy;
    a.fz[s] = fz;
    a.pos[o3] = x.x;
    a.pos[o3 + 1] = x.y;
    a.pos[o3 + 2] = x.z;
    if (o == a.o0) a.flags[a.unwrapped_flag] = 0;
    if (a.always_rebuild) {
        if (o == 0) a.flags[FLAG_REBUILD] = 1;
    } else {
        double dx = x.x - bx, dy = x.y - by, dz = x.z - bz;
        min_image<ORTHO>(a.box, dx, dy, dz);
        if (!(norm2(dx, dy, dz) <= a.half_skin2)) a.flags[FLAG_REBUILD] = 1;
    }
}

struct StoreRangeArgs {
    int o0, o1;
    const int *slot_of_id;
    const double *vx, *vy, *vz, *fx, *fy, *fz;
    double *vel, *frc;  // AoS staging (device)
};

__global__ void __launch_bounds__(TPB) k_store_range(StoreRangeArgs a) {
    const int o = a.o0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= a.o1) return;
    const int s = a.slot_of_id[o];
    const size_t o3 = 3 * (size_t)o;
    a.vel[o3] = a.vx[s];
    a.vel[o3 + 1] = a.vy[s];
    a.vel[o3 + 2] = a.vz[s];
    a.frc[o3] = a.fx[s];
    a.frc[o3 + 1] = a.fy[s];
    a.frc[o3 + 2] = a.fz[s];
}

// Skin trigger on freshly uploaded positions: any |minimg(x - x_build)|^2 > (skin/2)^2 => rebuild.
template <bool ORTHO>
__global__ void __launch_bounds__(TPB) k_check_displacement(int n, const double4 *__restrict__ xt,
                                                            const double *__restrict__ xbx,
                                                            const double *__restrict__ xby,
                                                            const double *__restrict__ xbz, BoxDev box,
                                                            double half_skin2, int *flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 x = xt[i];
    double dx = x.x - xbx[i], dy = x.y - xby[i], dz = x.z - xbz[i];
    min_image<ORTHO>(box, dx, dy, dz);
    double r2 = norm2(dx, dy, dz);
    if (!(r2 <= half_skin2)) flags[FLAG_REBUILD] = 1;
}

// ------------------------------------------------------------------------------------------------
// Fused velocity-Verlet kernel.
//   KICK : second half of step k     v += ((a_t + a_tdt) * 0.5) * dt           potential.rs:28-30
//          + per-step observables    KE (properties.rs:17-24), tr(X F^T) (properties.rs:49-51)
//   DRIFT: first half of step k+1    x += (v*dt) + ((a*0.5) * dt^2)            potential.rs:16-18
//          + wrap (potential.rs:20-22 -> simulation_box.rs:29-42) + skin trigger
//   a = F / m[type-1] is a true division (properties.rs:32-39).
// f = force at the current positions, g = force at the previous positions.
// ------------------------------------------------------------------------------------------------
struct VVArgs {
    int n;
    double4 *xt;
    float4 *xf;
    double *vx, *vy, *vz;
    const double *fx, *fy, *fz;  // F(t+dt) for KICK, F(t) for DRIFT
    const double *gx, *gy, *gz;  // F(t) for KICK
    const double *xbx, *xby, *xbz;
    const double *mass;  // per type
    BoxDev box;
    double dt, dt2, half_skin2;
    int always_rebuild;
    int *flags;
    double *partials;
    unsigned int *ticket;
    pisb_thermo *thermo;  // record of the step the KICK completes
    const double *vscale;  // NVT: thermostat scale exp(-dt/2 xi_1) read from device memory (null for NVE)
    double4 *xt_out;       // DRIFT: where x(t+dt) goes -- xt itself, or the other position buffer when the batch
    float4 *xf_out;        //        continues with k_force_vv launches (which alternate the two buffers)
    int unwrapped_out;     // DRIFT: flags[] word of the buffer behind xt_out (cleared: it now holds wrapped positions)
};

template <bool KICK, bool DRIFT, bool ORTHO>
__global__ void __launch_bounds__(TPB, 4) k_vv(VVArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[4] = {0.0, 0.0, 0.0, 0.0};  // ke, x*fx, y*fy, z*fz
    if (DRIFT && i == 0) a.flags[a.unwrapped_out] = 0;  // nobody reads this word while a drift is running
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        // every independent load first (the kernel is latency-bound otherwise: the mass lookup depends on x.w)
        double4 x = a.xt[i];
        double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        const double fx = a.fx[i], fy = a.fy[i], fz = a.fz[i];
        double gx = 0.0, gy = 0.0, gz = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
        if (KICK) {
            gx = a.gx[i];
            gy = a.gy[i];
            gz = a.gz[i];
        }
        if (DRIFT && !a.always_rebuild) {
            bx = a.xbx[i];
            by = a.xby[i];
            bz = a.xbz[i];
        }
        const double m = a.mass[type_of(x.w) - 1];
        const double ax = __ddiv_rn(fx, m), ay = __ddiv_rn(fy, m), az = __ddiv_rn(fz, m);
        if (KICK) {
            const double ox = __ddiv_rn(gx, m), oy = __ddiv_rn(gy, m), oz = __ddiv_rn(gz, m);
            vx = __dadd_rn(vx, __dmul_rn(__dmul_rn(__dadd_rn(ox, ax), 0.5), a.dt));
            vy = __dadd_rn(vy, __dmul_rn(__dmul_rn(__dadd_rn(oy, ay), 0.5), a.dt));
            vz = __dadd_rn(vz, __dmul_rn(__dmul_rn(__dadd_rn(oz, az), 0.5), a.dt));
            if (a.vscale) {  // verlet_step_nvt_nhc: velocities = &velocities * scale AFTER the NVE step (potential.rs:50)
                const double sc = *a.vscale;
                vx = __dmul_rn(vx, sc);
                vy = __dmul_rn(vy, sc);
                vz = __dmul_rn(vz, sc);
            }
            a.vx[i] = vx;
            a.vy[i] = vy;
            a.vz[i] = vz;
            red[0] = __dmul_rn(__dmul_rn(0.5, m), norm2(vx, vy, vz));
            red[1] = __dmul_rn(x.x, fx);
            red[2] = __dmul_rn(x.y, fy);
            red[3] = __dmul_rn(x.z, fz);
        }
        if (DRIFT) {
            if (!KICK && a.vscale) {  // ... and BEFORE it (potential.rs:45-46); stored, the kick starts from the scaled v
                const double sc = *a.vscale;
                vx = __dmul_rn(vx, sc);
                vy = __dmul_rn(vy, sc);
                vz = __dmul_rn(vz, sc);
                a.vx[i] = vx;
                a.vy[i] = vy;
                a.vz[i] = vz;
            }
            x.x = __dadd_rn(x.x, __dadd_rn(__dmul_rn(vx, a.dt), __dmul_rn(__dmul_rn(ax, 0.5), a.dt2)));
            x.y = __dadd_rn(x.y, __dadd_rn(__dmul_rn(vy, a.dt), __dmul_rn(__dmul_rn(ay, 0.5), a.dt2)));
            x.z = __dadd_rn(x.z, __dadd_rn(__dmul_rn(vz, a.dt), __dmul_rn(__dmul_rn(az, 0.5), a.dt2)));
            wrap_pos<ORTHO>(a.box, x.x, x.y, x.z);
            a.xt_out[i] = x;
            a.xf_out[i] = make_float4((float)x.x, (float)x.y, (float)x.z, __int_as_float(type_of(x.w)));
            if (a.always_rebuild) {
                if (i == 0) a.flags[FLAG_REBUILD] = 1;
            } else {
                double dx = x.x - bx, dy = x.y - by, dz = x.z - bz;
                min_image<ORTHO>(a.box, dx, dy, dz);
                if (!(norm2(dx, dy, dz) <= a.half_skin2)) a.flags[FLAG_REBUILD] = 1;
            }
        }
    }
    if (KICK) {
        pisb_thermo *th = a.thermo;
        double t3[3];
        block_reduce_finalize<4, TPB>(red, a.partials, a.ticket, [&](int q, double s) {
            if (q == 0) th->ke = s;
            else t3[q - 1] = s;
            if (q == 3) th->virial_ref = (t3[0] + t3[1]) + t3[2];
        });
    }
}

// ------------------------------------------------------------------------------------------------
// Nose-Hoover chain (src/ensemble/nvt.rs), kept on the device so an NVT batch needs no host round
// trip: one thread advances the chain from the kinetic energy the kick kernel just reduced.
// Quirks kept as written in the reference: xi is ASSIGNED in propagate_half_step (:72-80), eta never
// advances, the chain has 3 links (:108), the target temperature ramps linearly per step (:114-123).
// ------------------------------------------------------------------------------------------------
struct NhcDev {
    pisb_nhc c;
    double scale;          // exp(-0.5 dt xi[0]) of the current step
    double ke_last;        // kinetic energy the next first half step starts from (graph replays carry it on the device)
    long long step_index;  // index of the step in progress (graph replays: the ramp index advances on the device)
};

__global__ void k_nhc_half(NhcDev *nhc, const double *ke_ptr, long long n_atoms, double dt, int second_half,
                           long long step_index, long long total_steps, double *energy_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    pisb_nhc &c = nhc->c;
    const double KB = 0.0083144621;  // src/constants.rs:3
    const double ke = ke_ptr ? *ke_ptr : nhc->ke_last;  // null: carried on the device
    if (step_index < 0) step_index = nhc->step_index;   // negative: device-resident step counter
    // compute_forces (:59-69)
    c.g[0] = __dsub_rn(__dmul_rn(2.0, ke), __dmul_rn(__dmul_rn((double)(n_atoms * 3), KB), c.target_temperature));
    for (int j = 1; j < 3; ++j)
        c.g[j] = __dsub_rn(__dmul_rn(c.q[j - 1], __dmul_rn(c.xi[j - 1], c.xi[j - 1])), __dmul_rn(KB, c.target_temperature));
    // propagate_half_step (:72-80)
    c.xi[2] = __ddiv_rn(__dmul_rn(__dmul_rn(0.5, dt), c.g[2]), c.q[2]);
    for (int l = 1; l >= 0; --l)
        c.xi[l] = __dmul_rn(__ddiv_rn(__dmul_rn(__dmul_rn(0.5, dt), c.g[l]), c.q[l]), exp(__dmul_rn(__dmul_rn(-0.25, dt), c.xi[l + 1])));
    if (!second_half) {
        nhc->scale = exp(__dmul_rn(__dmul_rn(-0.5, dt), c.xi[0]));  // potential.rs:45
    } else {
        nhc->ke_last = ke;
        nhc->step_index = step_index + 1;
        // Simulation::step: calculate_target_temperature(i, steps) (simulation.rs:55; nvt.rs:114-123)
        c.target_temperature = __dadd_rn(c.start_temperature,
                                         __dmul_rn(__ddiv_rn(__dsub_rn(c.end_temperature, c.start_temperature), (double)total_steps),
                                                   (double)step_index));
        if (energy_out) {  // nhc.kinetic_energy() + nhc.potential_energy(n) for the Hamiltonian (simulation.rs:101-104)
            double tke = 0.0;
            for (int i = 0; i < 3; ++i) tke = __dadd_rn(tke, __dmul_rn(__dmul_rn(0.5, c.q[i]), __dmul_rn(c.xi[i], c.xi[i])));
            double tpe = __dmul_rn(__dmul_rn(__dmul_rn((double)(n_atoms * 3), KB), c.target_temperature), c.eta[0]);
            for (int i = 1; i < 3; ++i) tpe = __dadd_rn(tpe, __dmul_rn(__dmul_rn(KB, c.target_temperature), c.eta[i]));
            *energy_out = __dadd_rn(tke, tpe);
        }
    }
}

// KE / virial_ref of the current state without touching it (pisb_thermo_now).
__global__ void __launch_bounds__(TPB) k_observe(int n, const double4 *__restrict__ xt,
                                                 const double *__restrict__ vx, const double *__restrict__ vy,
                                                 const double *__restrict__ vz, const double *__restrict__ fx,
                                                 const double *__restrict__ fy, const double *__restrict__ fz,
                                                 const double *__restrict__ mass, double *partials,
                                                 unsigned int *ticket, pisb_thermo *th) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < n && !is_ghost(xt[i].w)) {
        double4 x = xt[i];
        const double m = mass[type_of(x.w) - 1];
        red[0] = __dmul_rn(__dmul_rn(0.5, m), norm2(vx[i], vy[i], vz[i]));
        red[1] = __dmul_rn(x.x, fx[i]);
        red[2] = __dmul_rn(x.y, fy[i]);
        red[3] = __dmul_rn(x.z, fz[i]);
    }
    double t3[3];
    block_reduce_finalize<4, TPB>(red, partials, ticket, [&](int q, double s) {
        if (q == 0) th->ke = s;
        else t3[q - 1] = s;
        if (q == 3) th->virial_ref = (t3[0] + t3[1]) + t3[2];
    });
}

// ------------------------------------------------------------------------------------------------
// Cell binning (replaces Atoms::rcut_cells, src/atoms/neighbour_list.rs:9-42).  The cell index is
// (cz*ny + cy)*nx + cx like cell_index (:61-63).  Unlike the reference, the fractional coordinate
// is wrapped before binning so that unwrapped step-0 inputs land in their periodic cell (the
// reference saturates negatives to cell 0 and then misses pairs: SURVEY appendix A.2).
// All rebuild-chain kernels exit immediately unless flags[FLAG_REBUILD] is set: the chain is
// launched every step and the skin trigger decides on the device, with no host round trip.
// ------------------------------------------------------------------------------------------------
template <bool ORTHO>
__global__ void __launch_bounds__(TPB) k_bin(int n, const double4 *__restrict__ xt, BoxDev box, Grid g,
                                             int *__restrict__ cell_of, int *__restrict__ cell_count,
                                             const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    // grid-stride: the chain is launched every step with a CAPPED grid, so the common no-rebuild case costs a few
    // microseconds instead of draining a 4M-thread grid of early exits
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double4 x = xt[i];
        int c[3];
        cell_coords<ORTHO>(box, g, x.x, x.y, x.z, c);
        int cell = (c[2] * g.n[1] + c[1]) * g.n[0] + c[0];
        // sentinel bucket (multi-GPU rebuilds): every stale ghost and every leaver lands on ONE counter -- 0.4M same-address
        // atomics per rebuild of a 4M-atom brick (bin 0.065 against 0.017 ms per step) unless a warp sends one
        const bool dead = is_dead(x.w);
        const unsigned act = __activemask();
        const unsigned md = __ballot_sync(act, dead);
        if (dead) {
            cell = g.ncell;
            if ((threadIdx.x & 31) == __ffs(md) - 1) atomicAdd(&cell_count[cell], __popc(md));
        } else {
            atomicAdd(&cell_count[cell], 1);
        }
        cell_of[i] = cell;
    }
}

// Exclusive scan of cell_count -> cell_start, three phases over tiles of SCAN_TILE entries.
constexpr int SCAN_TPB = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_TPB * SCAN_IPT;

__device__ __forceinline__ int block_exclusive_scan_int(int v, int &total) {
    __shared__ int s_w[SCAN_TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_TPB / 32 ? s_w[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < SCAN_TPB / 32) s_w[lane] = w;
    }
    __syncthreads();
    int base = warp > 0 ? s_w[warp - 1] : 0;
    total = s_w[SCAN_TPB / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_tiles(int ncell, const int *__restrict__ cell_count,
                                                         int *__restrict__ tile_sum, const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k)
        if (base + k < ncell) s += cell_count[base + k];
    int total;
    block_exclusive_scan_int(s, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_sums(int ntiles, int *__restrict__ tile_sum, const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    int carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_TPB) {
        int idx = base + threadIdx.x;
        int v = idx < ntiles ? tile_sum[idx] : 0;
        int total;
        int ex = block_exclusive_scan_int(v, total);
        if (idx < ntiles) tile_sum[idx] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_apply(int ncell, int n_atoms, int *__restrict__ cell_count,
                                                         const int *__restrict__ tile_sum,
                                                         int *__restrict__ cell_start, const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int v[SCAN_IPT];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        v[k] = base + k < ncell ? cell_count[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan_int(s, total) + tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < ncell) {
            cell_start[base + k] = ex;
            cell_count[base + k] = 0;  // reused as the fill cursor by k_fill
        }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cell_start[ncell] = n_atoms;
}

// Scatter slot indices into their cell segment (arrival order, fixed up by k_sort_cells).
__global__ void __launch_bounds__(TPB) k_fill(int n, int ncell, const int *__restrict__ cell_of,
                                              const int *__restrict__ cell_start, int *__restrict__ cell_fill,
                                              int *__restrict__ order, const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cell_of[i];
        // the sentinel bucket (cell index ncell, dead slots of a multi-GPU rebuild) again gets one atomic per warp
        const bool dead = c == ncell;
        const unsigned act = __activemask();
        const unsigned md = __ballot_sync(act, dead);
        int p;
        if (dead) {
            const int lane = threadIdx.x & 31, leader = __ffs(md) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cell_fill[c], __popc(md));
            base = __shfl_sync(md, base, leader);
            p = cell_start[c] + base + __popc(md & ((1u << lane) - 1u));
        } else {
            p = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        }
        order[p] = i;
    }
}

// Within each cell order atoms by ORIGINAL id: deterministic regardless of atomic arrival order,
// and the same within-cell order as the reference's push loop (ascending atom index,
// src/atoms/neighbour_list.rs:17,38).  One thread per cell; cells hold ~16 atoms.
__global__ void __launch_bounds__(TPB) k_sort_cells(int ncell, const int *__restrict__ cell_start,
                                                    int *__restrict__ order, const int *__restrict__ id,
                                                    const int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    for (int a = b + 1; a < e; ++a) {
        int oa = order[a];
        int ka = id[oa];
        int p = a - 1;
        while (p >= b) {
            int op = order[p];
            if (id[op] <= ka) break;
            order[p + 1] = op;
            --p;
        }
        order[p + 1] = oa;
    }
}

// Permute every per-atom array into the new cell-sorted order (gather into scratch), and record
// the build positions.  k_copy_back then moves scratch into the primary arrays, so the host-side
// pointers never depend on whether the device decided to rebuild.
struct PermArgs {
    int n;
    const int *order;
    const double4 *xt;
    const double *vx, *vy, *vz, *fx, *fy, *fz;
    const int *id;
    double4 *s_xt;
    double *s_vx, *s_vy, *s_vz, *s_fx, *s_fy, *s_fz;
    int *s_id;
    const int *flags;
};

__global__ void __launch_bounds__(TPB) k_permute(PermArgs a) {
    if (a.flags[FLAG_REBUILD] == 0) return;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.n; p += gridDim.x * blockDim.x) {
        int o = a.order[p];
        a.s_xt[p] = a.xt[o];
        a.s_vx[p] = a.vx[o];
        a.s_vy[p] = a.vy[o];
        a.s_vz[p] = a.vz[o];
        a.s_fx[p] = a.fx[o];
        a.s_fy[p] = a.fy[o];
        a.s_fz[p] = a.fz[o];
        a.s_id[p] = a.id[o];
    }
}

struct CopyBackArgs {
    int n;
    const double4 *s_xt;
    const double *s_vx, *s_vy, *s_vz, *s_fx, *s_fy, *s_fz;
    const int *s_id;
    double4 *xt;
    double *vx, *vy, *vz, *fx, *fy, *fz;
    int *id;
    int *slot_of_id;
    double *xbx, *xby, *xbz;
    const int *flags;
    float4 *xf;
    BoxDev box;
    float *xp;  // pair-packed FP32 positions for k_build_list_v3 (null: not written)
};

__global__ void __launch_bounds__(TPB) k_copy_back(CopyBackArgs a) {
    if (a.flags[FLAG_REBUILD] == 0) return;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.n; p += gridDim.x * blockDim.x) {
        double4 x = a.s_xt[p];
        a.xt[p] = x;
        const float4 xf = make_xf(a.box, x);
        a.xf[p] = xf;
        if (a.xp) {  // slots 2q, 2q+1 share one 32-byte record {x0,x1, y0,y1, z0,z1, w0,w1}
            float *rec = a.xp + (size_t)(p >> 1) * 8 + (p & 1);
            rec[0] = xf.x;
            rec[2] = xf.y;
            rec[4] = xf.z;
            rec[6] = xf.w;
        }
        a.xbx[p] = x.x;
        a.xby[p] = x.y;
        a.xbz[p] = x.z;
        a.vx[p] = a.s_vx[p];
        a.vy[p] = a.s_vy[p];
        a.vz[p] = a.s_vz[p];
        a.fx[p] = a.s_fx[p];
        a.fy[p] = a.s_fy[p];
        a.fz[p] = a.s_fz[p];
        int o = a.s_id[p];
        a.id[p] = o;
        if (a.slot_of_id) a.slot_of_id[o] = p;  // null in multi-GPU mode (ids are global there)
    }
}

// ------------------------------------------------------------------------------------------------
// Verlet-list build: FULL list, one thread per atom over the 27-cell stencil.
// Semantics: LJVPBuildListManager::build_neighbour_list (src/potentials/lennard_jones.rs:345-415)
// with rcut := rcut + skin: skip i == j, minimum image, `|rij| > rcut -> skip` (inclusive cutoff).
// Atoms of one cell are consecutive slots, so the lanes of a warp that share a cell walk the same
// candidate sequence (broadcast loads) and the transposed list rows are written coalesced.
// ------------------------------------------------------------------------------------------------
struct BuildArgs {
    int n, npad, kcap;
    const double4 *xt;
    const int *cell_start;
    BoxDev box;
    Grid g;
    PairDev pair0;           // single-type fast path
    const PairDev *table;    // n_types x n_types (MULTI)
    int n_types;
    int *nbr;
    int *nnbr;
    int *flags;
};

template <bool ORTHO, bool MULTI>
__global__ void __launch_bounds__(TPB_FORCE) k_build_list(BuildArgs a) {
    if (a.flags[FLAG_REBUILD] == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (i < a.n && is_ghost(a.xt[i].w)) a.nnbr[i] = 0;
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double4 xi = a.xt[i];
        const int ti = MULTI ? type_of(xi.w) : 1;
        int c[3];
        cell_coords<ORTHO>(a.box, a.g, xi.x, xi.y, xi.z, c);
        for (int dz = a.g.lo[2]; dz <= a.g.hi[2]; ++dz) {
            int cz = c[2] + dz;
            if (cz < 0 || cz >= a.g.n[2]) {
                if (!a.box.pbc[2] || a.g.local[2]) continue;
                cz += cz < 0 ? a.g.n[2] : -a.g.n[2];
            }
            for (int dy = a.g.lo[1]; dy <= a.g.hi[1]; ++dy) {
                int cy = c[1] + dy;
                if (cy < 0 || cy >= a.g.n[1]) {
                    if (!a.box.pbc[1] || a.g.local[1]) continue;
                    cy += cy < 0 ? a.g.n[1] : -a.g.n[1];
                }
                for (int dx = a.g.lo[0]; dx <= a.g.hi[0]; ++dx) {
                    int cx = c[0] + dx;
                    if (cx < 0 || cx >= a.g.n[0]) {
                        if (!a.box.pbc[0] || a.g.local[0]) continue;
                        cx += cx < 0 ? a.g.n[0] : -a.g.n[0];
                    }
                    const int cell = (cz * a.g.n[1] + cy) * a.g.n[0] + cx;
                    const int jb = __ldg(&a.cell_start[cell]), je = __ldg(&a.cell_start[cell + 1]);
                    for (int j = jb; j < je; ++j) {
                        if (j == i) continue;
                        const double4 xj = ldg_d4(&a.xt[j]);
                        double ddx = __dsub_rn(xj.x, xi.x), ddy = __dsub_rn(xj.y, xi.y), ddz = __dsub_rn(xj.z, xi.z);
                        min_image<ORTHO>(a.box, ddx, ddy, ddz);
                        const double r2 = norm2(ddx, ddy, ddz);
                        double t_list;
                        if (MULTI) {
                            const int tj = type_of(xj.w);
                            const int lo = min(ti, tj), hi = max(ti, tj);
                            const PairDev *p = &a.table[(lo - 1) * a.n_types + (hi - 1)];
                            if (!p->present) continue;
                            t_list = p->t_list;
                        } else {
                            t_list = a.pair0.t_list;
                        }
                        if (r2 > t_list) continue;
                        if (cnt < a.kcap) a.nbr[nbr_at(cnt, i, a.npad)] = j;
                        ++cnt;
                    }
                }
            }
        }
        nbr_close_row(a.nbr, i, a.npad, cnt, a.kcap);
        a.nnbr[i] = cnt < a.kcap ? cnt : a.kcap;
    }
    // record the largest list (overflow is detected by the host as max > kcap)
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&a.flags[FLAG_MAXNBR], m);
}

__global__ void k_finish_rebuild(int *flags) {
    if (flags[FLAG_REBUILD] == 0) return;
    flags[FLAG_REBUILD] = 0;
    flags[FLAG_NBUILDS] += 1;
}

// ------------------------------------------------------------------------------------------------
// LJ force / energy / virial over the Verlet list: one thread per atom, full list, no atomics.
// Replaces LJVOffsetManager::compute_potential (src/potentials/lennard_jones.rs:186-244) through
// the list consumer's form (:419-455): F_i += -f_ij per listed j inside rcut, PE = sum(u)/2.
// Per-pair terms are bit-identical to the reference's; only the summation order differs.
// ------------------------------------------------------------------------------------------------
struct ForceArgs {
    int n, npad;
    const double4 *xt;
    const int *nbr;
    const int *nnbr;
    BoxDev box;
    PairDev pair0;
    const PairDev *table;
    int n_types;
    const double *ax, *ay, *az;  // accumulate source (may alias the outputs) or null
    double *fx, *fy, *fz;        // output
    double *partials;
    unsigned int *ticket;
    pisb_thermo *thermo;
};

// Accumulation of one in-range pair into the per-atom sums.  u and fs are the reference's values bit for bit
// (lj_pair); the force sum F_i -= fs * r_ij (lennard_jones.rs:230-232) is accumulated with a fused multiply-add -- one
// rounding per term instead of two, 3 of 38 FP64 issue slots per listed pair saved on the pipe that bounds the kernel.
// Every force kernel uses this macro, so the variants stay bit-identical to one another.
#define PISB_ACCUM(dx, dy, dz, r2, u, fs) \
    do {                                  \
        fx = fma(-(fs), dx, fx);          \
        fy = fma(-(fs), dy, fy);          \
        fz = fma(-(fs), dz, fz);          \
        pe = __dadd_rn(pe, u);            \
        vir = fma(fs, r2, vir);           \
    } while (0)

template <bool ORTHO, bool MULTI>
__global__ void __launch_bounds__(TPB_FORCE) k_force(ForceArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[2] = {0.0, 0.0};  // sum u, sum fs*r2
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double4 xi = a.xt[i];
        const int ti = MULTI ? type_of(xi.w) : 1;
        const int nn = a.nnbr[i];
        double fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0, vir = 0.0;
        int jn = nn > 0 ? __ldg(a.nbr + nbr_at(0, i, a.npad)) : 0;
        for (int k = 0; k < nn; ++k) {
            const int j = jn;
            if (k + 1 < nn) jn = __ldg(a.nbr + nbr_at(k + 1, i, a.npad));
            const double4 xj = ldg_d4(&a.xt[j]);
            double dx = __dsub_rn(xj.x, xi.x), dy = __dsub_rn(xj.y, xi.y), dz = __dsub_rn(xj.z, xi.z);
            min_image<ORTHO>(a.box, dx, dy, dz);
            const double r2 = norm2(dx, dy, dz);
            PairDev p;
            if (MULTI) {
                const int tj = type_of(xj.w);
                const int lo = min(ti, tj), hi = max(ti, tj);
                p = a.table[(lo - 1) * a.n_types + (hi - 1)];
                if (!p.present) continue;
            } else {
                p = a.pair0;
            }
            if (r2 > p.t_rc) continue;
            double u, fs;
            lj_pair(p, r2, u, fs);
            PISB_ACCUM(dx, dy, dz, r2, u, fs);
        }
        if (a.ax) {
            fx += a.ax[i];
            fy += a.ay[i];
            fz += a.az[i];
        }
        a.fx[i] = fx;
        a.fy[i] = fy;
        a.fz[i] = fz;
        red[0] = pe;
        red[1] = vir;
    }
    pisb_thermo *th = a.thermo;
    block_reduce_finalize<2, TPB_FORCE>(red, a.partials, a.ticket, [&](int q, double s) {
        if (q == 0) th->pe = s / 2.0;
        else th->virial_pair = s / 2.0;
    });
}

// ================================================================================================
// v2 kernels (orthorhombic boxes): FP32 pre-filter + per-thread compaction + exact FP64 physics.
//
// B200 has 128 FP32 lanes but only 64 FP64 lanes per SM, FRND/F2F run at quarter rate, and the
// exact cutoff test costs ~33 FP64 issue slots per listed pair.  So the in/out decision is made on
// the otherwise idle FP32 pipe from a float4 copy of the (wrapped) positions:
//     r2f > hi  -> certainly outside        r2f < lo -> certainly inside
//     otherwise (a relative guard band of a few 1e-4 around the threshold, sized on the host from
//     the FP32 error bound (35 L/rc + 8) 2^-24) -> the exact FP64 reference predicate decides.
// The decision is therefore bit-identical to the reference's `rij.norm() > rcut` for every pair.
// Pairs found inside are compacted into a per-thread shared-memory queue and evaluated in a second,
// divergence-free FP64 loop with the reference's operation order (per-pair terms stay bit-identical;
// rint-by-magic-constant replaces round(): ties (|s| = 0.5) only occur at |d| = L/2 >= rc + skin,
// i.e. never for an in-range pair).  Warps whose atoms all sit farther than rc+2 skin from every
// periodic face skip the image search (round(s) == 0 exactly).
// ================================================================================================
struct BoxF {
    float L[3], invL[3];
    float margin;  // rc_list + skin: atoms farther than this from every face need no image search
    int pbc[3];
};

struct PairF {
    float lo_rc, hi_rc, lo_list, hi_list;
};

constexpr int QCAP = 64;  // compaction queue entries per thread

__device__ __forceinline__ float rint_magic_f(float s) { return __fsub_rn(__fadd_rn(s, 12582912.0f), 12582912.0f); }
__device__ __forceinline__ double rint_magic_d(double s) {
    return __dsub_rn(__dadd_rn(s, 6755399441055744.0), 6755399441055744.0);
}

template <bool IMAGE>
__device__ __forceinline__ float r2_f32(const BoxF &b, const float4 &xi, const float4 &xj) {
    float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z;
    if (IMAGE) {
        if (b.pbc[0]) dx = fmaf(-rint_magic_f(dx * b.invL[0]), b.L[0], dx);
        if (b.pbc[1]) dy = fmaf(-rint_magic_f(dy * b.invL[1]), b.L[1], dy);
        if (b.pbc[2]) dz = fmaf(-rint_magic_f(dz * b.invL[2]), b.L[2], dz);
    }
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// exact reference r2 (orthorhombic), true round(): used in the guard band
__device__ __forceinline__ double r2_exact_ortho(const BoxDev &b, const double4 &xi, const double4 &xj) {
    double dx = __dsub_rn(xj.x, xi.x), dy = __dsub_rn(xj.y, xi.y), dz = __dsub_rn(xj.z, xi.z);
    min_image<true>(b, dx, dy, dz);
    return norm2(dx, dy, dz);
}

// reference displacement for a pair known to be in range (no ties possible): magic rint
template <bool IMAGE>
__device__ __forceinline__ void disp_inrange_ortho(const BoxDev &b, const double4 &xi, const double4 &xj,
                                                   double &dx, double &dy, double &dz) {
    dx = __dsub_rn(xj.x, xi.x);
    dy = __dsub_rn(xj.y, xi.y);
    dz = __dsub_rn(xj.z, xi.z);
    double sx = __dmul_rn(b.hinv[0], dx), sy = __dmul_rn(b.hinv[4], dy), sz = __dmul_rn(b.hinv[8], dz);
    if (IMAGE) {
        if (b.pbc[0]) sx = __dsub_rn(sx, rint_magic_d(sx));
        if (b.pbc[1]) sy = __dsub_rn(sy, rint_magic_d(sy));
        if (b.pbc[2]) sz = __dsub_rn(sz, rint_magic_d(sz));
    }
    dx = __dmul_rn(b.h[0], sx);
    dy = __dmul_rn(b.h[4], sy);
    dz = __dmul_rn(b.h[8], sz);
}

__device__ __forceinline__ bool is_interior(const BoxF &b, const float4 &x) {
    bool in = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float c = d == 0 ? x.x : (d == 1 ? x.y : x.z);
        if (b.pbc[d]) in = in && (c >= b.margin) && (c <= b.L[d] - b.margin);
    }
    return in;
}

struct Force2Args {
    int n, npad;
    const double4 *xt;
    const float4 *xf;
    const int *nbr;
    const int *nnbr;
    BoxDev box;
    BoxF boxf;
    PairDev pair0;
    PairF pairf0;
    const PairDev *table;
    const PairF *tablef;
    int n_types;
    const double *ax, *ay, *az;
    double *fx, *fy, *fz;
    double *partials;
    unsigned int *ticket;
    pisb_thermo *thermo;
    const int *skip_flag;  // multi-GPU: the launch is speculative and returns at once if *skip_flag != 0 (a rebuild comes first)
    const int *unwrapped;  // FLAG_UNWRAPPED word of the position buffer xt: non-zero => no interior shortcut
};

// One in-range pair, reference operation order (bit-identical per-pair terms).
template <bool MULTI, bool IMAGE>
__device__ __forceinline__ void pair_terms(const BoxDev &box, const PairDev &pair0, const PairDev *table, int n_types,
                                           int ti, const double4 &xi, const double4 &xj, double &dx, double &dy,
                                           double &dz, double &r2, double &u, double &fs) {
    disp_inrange_ortho<IMAGE>(box, xi, xj, dx, dy, dz);
    r2 = norm2(dx, dy, dz);
    if (MULTI) {
        const int tj = type_of(xj.w);
        const PairDev p = table[(min(ti, tj) - 1) * n_types + (max(ti, tj) - 1)];
        lj_pair(p, r2, u, fs);
    } else {
        lj_pair(pair0, r2, u, fs);
    }
}

constexpr int FU = 4;  // phase-1 unroll: independent index loads + gathers in flight per thread

template <bool MULTI, bool IMAGE>
__device__ __forceinline__ void force2_body(const Force2Args &a, int i, int (*q)[TPB_FORCE], double &fx, double &fy,
                                            double &fz, double &pe, double &vir) {
    const double4 xi = a.xt[i];
    const float4 xif = a.xf[i];
    const int ti = MULTI ? xf_type(xif) : 1;
    const int nn = a.nnbr[i];
    const int tid = threadIdx.x;
    int k = 0;
    while (true) {
        // ---- phase 1: FP32 filter, FU neighbours at a time; compact the in-range ones ----
        int cnt = 0;
        while (k < nn && cnt <= QCAP - FU) {
            int j[FU];
            float4 xjf[FU];
            {
                const int4 t = __ldg(reinterpret_cast<const int4 *>(a.nbr) + ((size_t)(k >> 2) * a.npad + i));
                j[0] = t.x;
                j[1] = (k + 1 < nn) ? t.y : i;
                j[2] = (k + 2 < nn) ? t.z : i;
                j[3] = (k + 3 < nn) ? t.w : i;
            }
#pragma unroll
            for (int u = 0; u < FU; ++u) xjf[u] = __ldg(&a.xf[j[u]]);
#pragma unroll
            for (int u = 0; u < FU; ++u) {
                const float r2f = r2_f32<IMAGE>(a.boxf, xif, xjf[u]);
                PairF pf;
                int pidx = 0;
                if (MULTI) {
                    const int tj = xf_type(xjf[u]);
                    pidx = (min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1);
                    pf = a.tablef[pidx];
                } else {
                    pf = a.pairf0;
                }
                bool in = r2f < pf.lo_rc;  // lo_rc < 0 encodes "pair absent"
                if (!in && r2f <= pf.hi_rc) {
                    const double4 xj = ldg_d4(&a.xt[j[u]]);
                    const double t_rc = MULTI ? a.table[pidx].t_rc : a.pair0.t_rc;
                    in = !(r2_exact_ortho(a.box, xi, xj) > t_rc);
                }
                if (in && (k + u < nn)) {
                    q[cnt][tid] = j[u];
                    ++cnt;
                }
            }
            k += FU;
        }
        // ---- phase 2: exact FP64 pair terms, two independent chains per iteration ----
        int c = 0;
        for (; c + 1 < cnt; c += 2) {
            const int j0 = q[c][tid], j1 = q[c + 1][tid];
            const double4 x0 = ldg_d4(&a.xt[j0]);
            const double4 x1 = ldg_d4(&a.xt[j1]);
            double dx0, dy0, dz0, r20, u0, fs0, dx1, dy1, dz1, r21, u1, fs1;
            pair_terms<MULTI, IMAGE>(a.box, a.pair0, a.table, a.n_types, ti, xi, x0, dx0, dy0, dz0, r20, u0, fs0);
            pair_terms<MULTI, IMAGE>(a.box, a.pair0, a.table, a.n_types, ti, xi, x1, dx1, dy1, dz1, r21, u1, fs1);
            PISB_ACCUM(dx0, dy0, dz0, r20, u0, fs0);
            PISB_ACCUM(dx1, dy1, dz1, r21, u1, fs1);
        }
        if (c < cnt) {
            const double4 x0 = ldg_d4(&a.xt[q[c][tid]]);
            double dx0, dy0, dz0, r20, u0, fs0;
            pair_terms<MULTI, IMAGE>(a.box, a.pair0, a.table, a.n_types, ti, xi, x0, dx0, dy0, dz0, r20, u0, fs0);
            PISB_ACCUM(dx0, dy0, dz0, r20, u0, fs0);
        }
        if (k >= nn) break;
    }
}

// ------------------------------------------------------------------------------------------------
// Lean force loop (the default kernels: k_force_v3, k_force_vv, k_force_split).  Arithmetic: pisb_device.cuh "Lean pair
// arithmetic".  The in/out decision stays the reference's: the squared distance computed the short way is within
// band = (4 L/rc + 32) 2^-51 (relative) of the reference's r2 = |h (h_inv d - round(h_inv d))|^2 -- the round trip
// through the fractional coordinate costs a relative error of a few ulp of L/|d| -- so outside [t_lo, t_hi] both
// agree, and inside it (one pair in ~10^11) the reference-order predicate is evaluated, out of line.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ bool exact_r2_gt(double h0, double h1, double h2, double i0, double i1, double i2, double xix, double xiy,
                                         double xiz, double xjx, double xjy, double xjz, double t) {
    double sx = __dmul_rn(i0, __dsub_rn(xjx, xix)), sy = __dmul_rn(i1, __dsub_rn(xjy, xiy)), sz = __dmul_rn(i2, __dsub_rn(xjz, xiz));
    sx = __dsub_rn(sx, round(sx));
    sy = __dsub_rn(sy, round(sy));
    sz = __dsub_rn(sz, round(sz));
    return norm2(__dmul_rn(h0, sx), __dmul_rn(h1, sy), __dmul_rn(h2, sz)) > t;
}

// per-atom sums of the lean loop; finish() applies the constant factors once
template <bool MULTI>
struct LeanAcc {
    double fx = 0.0, fy = 0.0, fz = 0.0;  // sum w d           (MULTI: sum c24 w d)
    double a = 0.0, b = 0.0;              // sum s12, sum s6   (MULTI: sum u, sum fs r2)
    int cnt = 0;                          // pairs in range (for the energy shift)
    __device__ __forceinline__ void pair(const PairDev &p, double dx, double dy, double dz, double r2) {
        const double y = rcp_newton(r2);
        const double s2 = p.sig2 * y;
        const double s6 = (s2 * s2) * s2;
        const double s12 = s6 * s6;
        const double t = fma(2.0, s12, -s6);
        if (MULTI) {
            const double w = (p.c24 * t) * y;
            fx = fma(w, dx, fx);
            fy = fma(w, dy, fy);
            fz = fma(w, dz, fz);
            a += fma(p.c4, s12 - s6, -p.ucut);
            b = fma(p.c24, t, b);
        } else {
            const double w = t * y;
            fx = fma(w, dx, fx);
            fy = fma(w, dy, fy);
            fz = fma(w, dz, fz);
            a += s12;
            b += s6;
            ++cnt;
        }
    }
    // F_i = -sum fs rij, PE contribution sum u, pair virial sum fs r2 (lennard_jones.rs:33-55, :230-232)
    __device__ __forceinline__ void finish(const PairDev &p0, double &ofx, double &ofy, double &ofz, double &pe, double &vir) const {
        if (MULTI) {
            ofx = -fx, ofy = -fy, ofz = -fz;
            pe = a;
            vir = b;
        } else {
            ofx = -p0.c24 * fx, ofy = -p0.c24 * fy, ofz = -p0.c24 * fz;
            pe = fma(p0.c4, a - b, -(double)cnt * p0.ucut);
            vir = p0.c24 * fma(2.0, a, -b);
        }
    }
};

// displacement xj - xi (minimum image when IMAGE) and its square, the short way
template <bool IMAGE>
__device__ __forceinline__ double lean_disp(const BoxDev &b, const double4 &xi, const double4 &xj, double &dx, double &dy, double &dz) {
    dx = xj.x - xi.x;
    dy = xj.y - xi.y;
    dz = xj.z - xi.z;
    if (IMAGE) {
        dx = fma(-rint_magic_d(dx * b.hinv[0]), b.h[0], dx);
        dy = fma(-rint_magic_d(dy * b.hinv[4]), b.h[4], dy);
        dz = fma(-rint_magic_d(dz * b.hinv[8]), b.h[8], dz);
    }
    return fma(dz, dz, fma(dy, dy, dx * dx));
}

// the reference's `rij.norm() > rcut -> skip` from the lean r2, reference-order predicate inside the guard band
__device__ __forceinline__ bool lean_in_range(const BoxDev &b, const PairDev &p, const double4 &xi, const double4 &xj, double r2) {
    if (le_bits(r2, p.t_lo)) return true;
    if (!le_bits(r2, p.t_hi)) return false;
    return !exact_r2_gt(b.h[0], b.h[4], b.h[8], b.hinv[0], b.hinv[4], b.hinv[8], xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, p.t_rc);
}

// An atom that met a guard-band pair in the fast loop (one pair in ~10^11 lands there) is redone here, start to finish,
// with the reference-order predicate available.  Out of line and called AFTER the loop: a call inside the hot loop would
// make every value live across it a spill (ptxas: 300 bytes of loop spills at the 64-register bound).
template <bool MULTI, bool IMAGE>
__device__ __noinline__ void force_atom_exact(const double4 *__restrict__ xt, const int *__restrict__ nbr, int npad, int i, int nn, int l, int stride,
                                              BoxDev box, PairDev pair0, const PairDev *__restrict__ table, int n_types, double *out5) {
    const double4 xi = xt[i];
    const int ti = MULTI ? type_of(xi.w) : 1;
    LeanAcc<MULTI> acc;
    for (int k0 = 4 * l; k0 < nn; k0 += 4 * stride)
        for (int k = k0; k < min(k0 + 4, nn); ++k) {
            const int j = nbr[nbr_at(k, i, npad)];
            const double4 xj = ldg_d4(&xt[j]);
            double dx, dy, dz;
            const double r2 = lean_disp<IMAGE>(box, xi, xj, dx, dy, dz);
            PairDev p = pair0;
            if (MULTI) {
                const int tj = type_of(xj.w);
                p = table[(min(ti, tj) - 1) * n_types + (max(ti, tj) - 1)];
                if (!p.present) continue;
            }
            if (lean_in_range(box, p, xi, xj, r2)) acc.pair(p, dx, dy, dz, r2);
        }
    acc.finish(pair0, out5[0], out5[1], out5[2], out5[3], out5[4]);
}

// One thread per atom: the list is read as one int4 per 4 neighbours and prefetched one tile ahead (the index stream
// comes from DRAM), four 256-bit gathers are in flight per thread, interior warps skip the image search.
template <bool MULTI, bool IMAGE>
__device__ __forceinline__ void force3_body(const Force2Args &a, int i, double &fx, double &fy, double &fz, double &pe,
                                            double &vir) {
    const double4 xi = a.xt[i];
    const int ti = MULTI ? type_of(xi.w) : 1;
    const int nn = a.nnbr[i];
    const int4 *tiles = reinterpret_cast<const int4 *>(a.nbr) + i;
    int4 cur = nn > 0 ? ldg_stream_i4(tiles) : make_int4(0, 0, 0, 0);
    LeanAcc<MULTI> acc;
    bool ambiguous = false;
    // (prefetch.global.L1 of the next tile's four records while the current tile is computed: 1.128 -> 1.509 ms per launch --
    // every extra L1TEX request is paid in full, profiles/r02_ab_prefetch_4M_43K.jsonl)
    // the loop carries one pointer and one count (k and nn as two live values cost the 64-register kernel a reload per tile)
    for (int rem = nn; rem > 0; rem -= 4) {
        int4 nxt = cur;
        tiles += a.npad;
        if (rem > 4) nxt = ldg_stream_i4(tiles);
        const int j[4] = {cur.x, cur.y, cur.z, cur.w};
        double4 xj[4];
        // a pad (the closed tail of the last tile) gathers the NaN record in front of slot 0 (reserve_positions): its r2 is
        // NaN and fails both range tests below -- one IMNMX, no predicate to keep alive
#pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = ldg_d4(&a.xt[max(j[u], -1)]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double dx, dy, dz;
            const double r2 = lean_disp<IMAGE>(a.box, xi, xj[u], dx, dy, dz);
            PairDev p = a.pair0;
            bool present = true;
            if (MULTI) {
                const int tj = type_of(xj[u].w);
                p = a.table[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)];
                present = p.present;
            }
            if (present) {
                if (le_bits(r2, p.t_lo)) acc.pair(p, dx, dy, dz, r2);
                else ambiguous |= le_bits(r2, p.t_hi);
            }
        }
        cur = nxt;
    }
    acc.finish(a.pair0, fx, fy, fz, pe, vir);
    if (ambiguous) {
        double o[5];
        force_atom_exact<MULTI, IMAGE>(a.xt, a.nbr, a.npad, i, nn, 0, 1, a.box, a.pair0, a.table, a.n_types, o);
        fx = o[0], fy = o[1], fz = o[2], pe = o[3], vir = o[4];
    }
}

// 64 registers, 32 warps per SM: measured best, profiles/r01_force_launch_config.txt
template <bool MULTI>
__global__ void __launch_bounds__(TPB_VV, 1024 / TPB_VV) k_force_v3(Force2Args a) {
    if (a.skip_flag && *a.skip_flag != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[2] = {0.0, 0.0};
    const bool active = i < a.n && !xf_is_ghost(a.xf[i]);
    bool interior = *a.unwrapped == 0;
    if (active) interior = interior && is_interior(a.boxf, a.xf[i]);
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    if (active) {
        double fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0, vir = 0.0;
        if (warp_interior) force3_body<MULTI, false>(a, i, fx, fy, fz, pe, vir);
        else force3_body<MULTI, true>(a, i, fx, fy, fz, pe, vir);
        if (a.ax) {
            fx += a.ax[i];
            fy += a.ay[i];
            fz += a.az[i];
        }
        a.fx[i] = fx;
        a.fy[i] = fy;
        a.fz[i] = fz;
        red[0] = pe;
        red[1] = vir;
    }
    pisb_thermo *th = a.thermo;
    block_reduce_finalize<2, TPB_VV>(red, a.partials, a.ticket, [&](int qq, double s) {
        if (qq == 0) th->pe = s / 2.0;
        else th->virial_pair = s / 2.0;
    });
}

// ------------------------------------------------------------------------------------------------
// k_force_vv<MULTI, DRIFT>: the force pass with the integrator in its epilogue -- ONE kernel per NVE step.
//   force  : k_force_v3's loop, unchanged (same pair terms, same summation order)
//   kick   : second half of step k      v += ((a_t + a_tdt) * 0.5) * dt      potential.rs:28-30, + KE, tr(X F^T)
//   drift  : first half of step k+1     x += (v*dt) + ((a*0.5) * dt^2), wrap, skin trigger   potential.rs:16-22
// The force pass is bound by the FP64 pipe and the L1TEX gather path and leaves ~80 % of the HBM bandwidth idle; the
// 200 B/atom the separate k_vv<kick,drift> launch streams (0.16 ms per step at 4M atoms) ride along for free here.
// Other threads still gather x(t) while this one already knows x(t+dt), so the drifted position goes to the OTHER
// position buffer (xt_out / xf_out); the host swaps the two after every drifting launch.
// Arithmetic of the epilogue is k_vv<true, DRIFT, true>'s, operation for operation.
// ------------------------------------------------------------------------------------------------
struct ForceVVArgs {
    Force2Args f;                // f.fx/fy/fz receive F(t+dt); f.ax must be null
    double *vx, *vy, *vz;
    const double *gx, *gy, *gz;  // F(t)
    const double *xbx, *xby, *xbz;
    const double *mass;
    double4 *xt_out;
    float4 *xf_out;
    double dt, dt2, half_skin2;
    int always_rebuild;
    int *flags;
    int unwrapped_out;  // flags[] word of the buffer behind xt_out (cleared by a drifting launch; never the word this launch reads)
};

// (Two scheduling experiments, profiles/r02_ab_force_block_sched_4M_43K.jsonl: blocks of 256 / 512 / 1024 threads -- 1.145 /
// 1.130 / 1.180 ms per launch against 1.149 with 128, so 512 it is; and an SM-aware work assignment that gives the resident
// blocks of one SM adjacent cell rows, so that they share neighbour records in L1 -- 3 % SLOWER at every block size.)
template <bool MULTI, bool DRIFT, bool BRICK, int NT = TPB_VV>
__global__ void __launch_bounds__(NT, 1024 / NT) k_force_vv(ForceVVArgs b) {
    const Force2Args &a = b.f;
    if (BRICK && a.skip_flag && *a.skip_flag != 0) return;  // speculative launch, a rebuild comes first
    const int i = blockIdx.x * NT + threadIdx.x;
    double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // pe, pair virial, ke, x*fx, y*fy, z*fz
    if (DRIFT && i == 0) b.flags[b.unwrapped_out] = 0;
    bool active = i < a.n;
    bool interior = *a.unwrapped == 0;
    if (active) {
        const float4 xfi = a.xf[i];  // one load serves both tests (written this way the force loop stays free of spills)
        if (BRICK && xf_is_ghost(xfi)) active = false;
        else interior = interior && is_interior(a.boxf, xfi);
    }
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    if (active) {
        double fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0, vir = 0.0;
        if (warp_interior) force3_body<MULTI, false>(a, i, fx, fy, fz, pe, vir);
        else force3_body<MULTI, true>(a, i, fx, fy, fz, pe, vir);
        a.fx[i] = fx;
        a.fy[i] = fy;
        a.fz[i] = fz;
        red[0] = pe;
        red[1] = vir;
        // ---- integrator epilogue ----
        double4 x = a.xt[i];  // L1/L2 hit: this thread read it at the top of the force loop
        double vx = b.vx[i], vy = b.vy[i], vz = b.vz[i];
        const double gx = b.gx[i], gy = b.gy[i], gz = b.gz[i];
        double bx = 0.0, by = 0.0, bz = 0.0;
        if (DRIFT && !b.always_rebuild) {
            bx = b.xbx[i];
            by = b.xby[i];
            bz = b.xbz[i];
        }
        const double m = b.mass[type_of(x.w) - 1];
        const double ax = __ddiv_rn(fx, m), ay = __ddiv_rn(fy, m), az = __ddiv_rn(fz, m);
        const double ox = __ddiv_rn(gx, m), oy = __ddiv_rn(gy, m), oz = __ddiv_rn(gz, m);
        vx = __dadd_rn(vx, __dmul_rn(__dmul_rn(__dadd_rn(ox, ax), 0.5), b.dt));
        vy = __dadd_rn(vy, __dmul_rn(__dmul_rn(__dadd_rn(oy, ay), 0.5), b.dt));
        vz = __dadd_rn(vz, __dmul_rn(__dmul_rn(__dadd_rn(oz, az), 0.5), b.dt));
        b.vx[i] = vx;
        b.vy[i] = vy;
        b.vz[i] = vz;
        red[2] = __dmul_rn(__dmul_rn(0.5, m), norm2(vx, vy, vz));
        red[3] = __dmul_rn(x.x, fx);
        red[4] = __dmul_rn(x.y, fy);
        red[5] = __dmul_rn(x.z, fz);
        if (DRIFT) {
            x.x = __dadd_rn(x.x, __dadd_rn(__dmul_rn(vx, b.dt), __dmul_rn(__dmul_rn(ax, 0.5), b.dt2)));
            x.y = __dadd_rn(x.y, __dadd_rn(__dmul_rn(vy, b.dt), __dmul_rn(__dmul_rn(ay, 0.5), b.dt2)));
            x.z = __dadd_rn(x.z, __dadd_rn(__dmul_rn(vz, b.dt), __dmul_rn(__dmul_rn(az, 0.5), b.dt2)));
            wrap_pos<true>(a.box, x.x, x.y, x.z);
            b.xt_out[i] = x;
            b.xf_out[i] = make_float4((float)x.x, (float)x.y, (float)x.z, __int_as_float(type_of(x.w)));
            if (b.always_rebuild) {
                if (i == 0) b.flags[FLAG_REBUILD] = 1;
            } else {
                double dx = x.x - bx, dy = x.y - by, dz = x.z - bz;
                min_image<true>(a.box, dx, dy, dz);
                if (!(norm2(dx, dy, dz) <= b.half_skin2)) b.flags[FLAG_REBUILD] = 1;
            }
        }
    }
    pisb_thermo *th = a.thermo;
    double t3[3];
    block_reduce_finalize<6, NT>(red, a.partials, a.ticket, [&](int q, double s) {
        if (q == 0) th->pe = s / 2.0;
        else if (q == 1) th->virial_pair = s / 2.0;
        else if (q == 2) th->ke = s;
        else t3[q - 3] = s;
        if (q == 5) th->virial_ref = (t3[0] + t3[1]) + t3[2];
    });
}

// ------------------------------------------------------------------------------------------------
// k_force_split<S>: S lanes per atom, for systems too small to fill 148 SMs with one thread per atom (the reference's
// own example has 4000 atoms: 32 blocks, one warp per scheduler, nothing to hide the gather + FP64 latency with;
// 28 us per launch with k_force_v3).  Lane l of an atom's group takes the K-tiles l, l+S, ... of its list row (same
// pair arithmetic, bit-identical pair terms); the S partial sums are combined by xor-shuffles in a fixed order.
// Only the summation order differs from k_force_v3.
// ------------------------------------------------------------------------------------------------
template <bool MULTI, bool IMAGE, int S>
__device__ __forceinline__ void force_split_body(const Force2Args &a, int i, int l, double &fx, double &fy, double &fz,
                                                 double &pe, double &vir) {
    const double4 xi = a.xt[i];
    const int ti = MULTI ? type_of(xi.w) : 1;
    const int nn = a.nnbr[i];
    const int4 *tiles = reinterpret_cast<const int4 *>(a.nbr) + i;
    LeanAcc<MULTI> acc;
    bool ambiguous = false;
    for (int k = 4 * l; k < nn; k += 4 * S) {
        const int4 cur = __ldg(tiles + (size_t)(k >> 2) * a.npad);
        int j[4] = {cur.x, cur.y, cur.z, cur.w};
        bool in[4];
        double4 xj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            in[u] = k + u < nn;
            if (!in[u]) j[u] = i;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = ldg_d4(&a.xt[j[u]]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double dx, dy, dz;
            const double r2 = lean_disp<IMAGE>(a.box, xi, xj[u], dx, dy, dz);
            PairDev p = a.pair0;
            if (MULTI) {
                const int tj = type_of(xj[u].w);
                p = a.table[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)];
                in[u] = in[u] && p.present;
            }
            if (in[u]) {
                if (le_bits(r2, p.t_lo)) acc.pair(p, dx, dy, dz, r2);
                else ambiguous |= le_bits(r2, p.t_hi);
            }
        }
    }
    acc.finish(a.pair0, fx, fy, fz, pe, vir);
    if (ambiguous) {  // this lane's share of the row again, with the reference-order predicate in the guard band
        double o[5];
        force_atom_exact<MULTI, IMAGE>(a.xt, a.nbr, a.npad, i, nn, l, S, a.box, a.pair0, a.table, a.n_types, o);
        fx = o[0], fy = o[1], fz = o[2], pe = o[3], vir = o[4];
    }
}

template <bool MULTI, int S>
__global__ void __launch_bounds__(TPB_FORCE) k_force_split(Force2Args a) {
    if (a.skip_flag && *a.skip_flag != 0) return;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt / S, l = gt % S;
    double red[2] = {0.0, 0.0};
    const bool active = i < a.n && !xf_is_ghost(a.xf[i]);
    bool interior = *a.unwrapped == 0;
    if (active) interior = interior && is_interior(a.boxf, a.xf[i]);
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    double fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0, vir = 0.0;
    if (active) {
        if (warp_interior) force_split_body<MULTI, false, S>(a, i, l, fx, fy, fz, pe, vir);
        else force_split_body<MULTI, true, S>(a, i, l, fx, fy, fz, pe, vir);
    }
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) {
        fx = __dadd_rn(fx, __shfl_xor_sync(0xffffffffu, fx, o));
        fy = __dadd_rn(fy, __shfl_xor_sync(0xffffffffu, fy, o));
        fz = __dadd_rn(fz, __shfl_xor_sync(0xffffffffu, fz, o));
        pe = __dadd_rn(pe, __shfl_xor_sync(0xffffffffu, pe, o));
        vir = __dadd_rn(vir, __shfl_xor_sync(0xffffffffu, vir, o));
    }
    if (active && l == 0) {
        if (a.ax) {
            fx += a.ax[i];
            fy += a.ay[i];
            fz += a.az[i];
        }
        a.fx[i] = fx;
        a.fy[i] = fy;
        a.fz[i] = fz;
        red[0] = pe;
        red[1] = vir;
    }
    pisb_thermo *th = a.thermo;
    block_reduce_finalize<2, TPB_FORCE>(red, a.partials, a.ticket, [&](int qq, double s) {
        if (qq == 0) th->pe = s / 2.0;
        else th->virial_pair = s / 2.0;
    });
}

// ------------------------------------------------------------------------------------------------
// k_force_q<MULTI, FUSED, DRIFT, BRICK>: FOUR LANES PER ATOM, lane l takes entry l of every K-tile.
//
// What binds the thread-per-atom loop after the FP64 diet is the L1TEX data pipe: one wavefront per distinct 128-byte
// line a gather instruction touches, and the k-th neighbours of 32 DIFFERENT atoms share few lines (ncu r01: ~25 lines per
// instruction, LSU data pipe 90 %).  The four entries of one K-tile of ONE atom, though, are consecutive acceptances of the
// build's slot-order scan, i.e. mostly adjacent slots of one cell run: 4 records = 1-2 lines.  With four lanes per atom a
// gather instruction covers the tiles of 8 atoms and touches ~17 lines instead of ~25 for the same 32 pairs (counted on
// real lists, tools/sim_lanes.py: 31 wavefronts per atom against 58; 8 lanes per atom: 30, 2 lanes: 39, two atoms per
// thread with shared entries: 48) -- the same pair evaluations, the same lists, the same arithmetic, ~half the L1 traffic.
// The tile layout [K/4][npad][4] is unchanged: the group's index load is one 16-byte tile.
// Four tiles per iteration keep four independent gathers in flight per lane; the next iteration's indices are loaded
// before the current gathers are consumed.  The 4 partial sums of an atom meet in an xor butterfly (fixed order, identical
// on all four lanes); then lane c < 3 owns COMPONENT c for the store and for the integrator epilogue
// (potential.rs:16-30: kick, KE, tr(X F^T), drift, wrap, skin trigger -- same operations as k_vv, spread over the lanes).
// ------------------------------------------------------------------------------------------------
constexpr int TPB_Q = 256;  // 64 atoms per block

__device__ __forceinline__ int ldg_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool MULTI, bool IMAGE>
__device__ __forceinline__ void force_q_body(const Force2Args &a, int i, int l, int nn, LeanAcc<MULTI> &acc, bool &ambiguous) {
    const double4 xi = a.xt[i];  // one address for the four lanes of the group
    const int ti = MULTI ? type_of(xi.w) : 1;
    const int *row = a.nbr + (size_t)i * 4 + l;  // entry 4t + l sits in tile t
    const size_t tstride = (size_t)a.npad * 4;
    int jn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) jn[u] = 4 * u + l < nn ? ldg_stream_i32(row + (size_t)u * tstride) : -1;
    for (int t = 0; 4 * t < nn; t += 4) {
        int j[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) j[u] = jn[u];
        if (4 * (t + 4) < nn) {
#pragma unroll
            for (int u = 0; u < 4; ++u) jn[u] = 4 * (t + 4 + u) + l < nn ? ldg_stream_i32(row + (size_t)(t + 4 + u) * tstride) : -1;
        }
        double4 xj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = ldg_d4(&a.xt[max(j[u], 0)]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double dx, dy, dz;
            const double r2 = lean_disp<IMAGE>(a.box, xi, xj[u], dx, dy, dz);
            PairDev p = a.pair0;
            bool in = j[u] >= 0;
            if (MULTI) {
                const int tj = type_of(xj[u].w);
                p = a.table[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)];
                in = in && p.present;
            }
            if (in) {
                if (le_bits(r2, p.t_lo)) acc.pair(p, dx, dy, dz, r2);
                else ambiguous |= le_bits(r2, p.t_hi);
            }
        }
    }
}

template <bool MULTI, bool FUSED, bool DRIFT, bool BRICK>
__global__ void __launch_bounds__(TPB_Q, 4) k_force_q(ForceVVArgs b) {
    const Force2Args &a = b.f;
    if (BRICK && a.skip_flag && *a.skip_flag != 0) return;  // speculative launch, a rebuild comes first
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt >> 2, l = gt & 3;
    double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // pe, pair virial, ke, x*fx, y*fy, z*fz
    if (FUSED && DRIFT && gt == 0) b.flags[b.unwrapped_out] = 0;
    bool active = i < a.n;
    bool interior = *a.unwrapped == 0;
    if (active) {
        const float4 xfi = a.xf[i];
        if (BRICK && xf_is_ghost(xfi)) active = false;
        else interior = interior && is_interior(a.boxf, xfi);
    }
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    LeanAcc<MULTI> acc;
    bool ambiguous = false;
    const int nn = active ? a.nnbr[i] : 0;
    if (active) {
        if (warp_interior) force_q_body<MULTI, false>(a, i, l, nn, acc, ambiguous);
        else force_q_body<MULTI, true>(a, i, l, nn, acc, ambiguous);
    }
    double fx, fy, fz, pe, vir;
    acc.finish(a.pair0, fx, fy, fz, pe, vir);
    // a guard-band pair anywhere in the row (one in ~10^11): the four lanes redo their shares (whole tiles l, l+4, ...)
    // with the reference-order predicate; out of line, after the loop
    ambiguous = __shfl_xor_sync(0xffffffffu, (int)ambiguous, 1) | (int)ambiguous;
    ambiguous = __shfl_xor_sync(0xffffffffu, (int)ambiguous, 2) | (int)ambiguous;
    if (ambiguous && active) {
        double o[5];
        if (warp_interior) force_atom_exact<MULTI, false>(a.xt, a.nbr, a.npad, i, nn, l, 4, a.box, a.pair0, a.table, a.n_types, o);
        else force_atom_exact<MULTI, true>(a.xt, a.nbr, a.npad, i, nn, l, 4, a.box, a.pair0, a.table, a.n_types, o);
        fx = o[0], fy = o[1], fz = o[2], pe = o[3], vir = o[4];
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
        vir += __shfl_xor_sync(0xffffffffu, vir, o);
    }
    // lane c < 3 owns component c from here on
    const int c = l;
    double fc = c == 0 ? fx : (c == 1 ? fy : fz);
    if (active) {
        if (l == 0) {
            red[0] = pe;
            red[1] = vir;
        }
        if (c < 3) {
            double *const fout = c == 0 ? a.fx : (c == 1 ? a.fy : a.fz);
            if (!FUSED && a.ax) fc += (c == 0 ? a.ax : (c == 1 ? a.ay : a.az))[i];
            fout[i] = fc;
        }
    }
    if (FUSED) {
        // ---- integrator epilogue, one component per lane (lane 3 carries the type word of the position record) ----
        double xc = 0.0, vc = 0.0, v2 = 0.0, d2 = 0.0;
        const double4 x4 = active ? a.xt[i] : make_double4(0.0, 0.0, 0.0, 0.0);
        if (active && c < 3) {
            xc = c == 0 ? x4.x : (c == 1 ? x4.y : x4.z);
            double *const vp = c == 0 ? b.vx : (c == 1 ? b.vy : b.vz);
            const double gc = (c == 0 ? b.gx : (c == 1 ? b.gy : b.gz))[i];
            vc = vp[i];
            const double m = b.mass[type_of(x4.w) - 1];
            const double ac = __ddiv_rn(fc, m), oc = __ddiv_rn(gc, m);
            vc = __dadd_rn(vc, __dmul_rn(__dmul_rn(__dadd_rn(oc, ac), 0.5), b.dt));
            vp[i] = vc;
            v2 = __dmul_rn(vc, vc);
            red[3 + c] = __dmul_rn(xc, fc);
            if (DRIFT) {
                double bc = 0.0;
                if (!b.always_rebuild) bc = (c == 0 ? b.xbx : (c == 1 ? b.xby : b.xbz))[i];
                xc = __dadd_rn(xc, __dadd_rn(__dmul_rn(vc, b.dt), __dmul_rn(__dmul_rn(ac, 0.5), b.dt2)));
                // wrap (simulation_box.rs:29-42, orthorhombic): s = h_inv x; s -= floor(s); x = h s
                const double hc = a.box.h[4 * c], ic = a.box.hinv[4 * c];
                double sc = __dmul_rn(ic, xc);
                sc = __dsub_rn(sc, floor(sc));
                xc = __dmul_rn(hc, sc);
                // skin trigger: minimum image of x - x_build (simulation_box.rs:17-27), squared
                double dc = xc - bc;
                double sd = __dmul_rn(ic, dc);
                sd = __dsub_rn(sd, round(sd));
                dc = __dmul_rn(hc, sd);
                d2 = __dmul_rn(dc, dc);
            }
        }
        // KE = 0.5 m ((vx^2 + vy^2) + vz^2) and |d|^2 = (dx^2 + dy^2) + dz^2 in the reference's association, on lane 0
        const unsigned full = 0xffffffffu;
        const int base = threadIdx.x & 28;  // lane 0 of this group within the warp
        const double v2y = __shfl_sync(full, v2, base + 1), v2z = __shfl_sync(full, v2, base + 2);
        const double d2y = __shfl_sync(full, d2, base + 1), d2z = __shfl_sync(full, d2, base + 2);
        if (active && l == 0) {
            const double m = b.mass[type_of(x4.w) - 1];
            red[2] = __dmul_rn(__dmul_rn(0.5, m), __dadd_rn(__dadd_rn(v2, v2y), v2z));
            if (DRIFT) {
                if (b.always_rebuild) {
                    if (i == 0) b.flags[FLAG_REBUILD] = 1;
                } else if (!(__dadd_rn(__dadd_rn(d2, d2y), d2z) <= b.half_skin2)) {
                    b.flags[FLAG_REBUILD] = 1;
                }
            }
        }
        if (DRIFT && active) {
            // position record {x, y, z, type bits} and its FP32 shadow: one 8-byte / 4-byte word per lane, 32 / 16 contiguous bytes per atom
            double *const xo = reinterpret_cast<double *>(b.xt_out + i);
            float *const fo = reinterpret_cast<float *>(b.xf_out + i);
            xo[c] = c < 3 ? xc : x4.w;
            fo[c] = c < 3 ? (float)xc : __int_as_float(type_of(x4.w));
        }
    }
    pisb_thermo *th = a.thermo;
    if (FUSED) {
        double t3[3];
        block_reduce_finalize<6, TPB_Q>(red, a.partials, a.ticket, [&](int q, double s) {
            if (q == 0) th->pe = s / 2.0;
            else if (q == 1) th->virial_pair = s / 2.0;
            else if (q == 2) th->ke = s;
            else t3[q - 3] = s;
            if (q == 5) th->virial_ref = (t3[0] + t3[1]) + t3[2];
        });
    } else {
        double r2[2] = {red[0], red[1]};
        block_reduce_finalize<2, TPB_Q>(r2, a.partials, a.ticket, [&](int q, double s) {
            if (q == 0) th->pe = s / 2.0;
            else th->virial_pair = s / 2.0;
        });
    }
}

template <bool MULTI>
__global__ void __launch_bounds__(TPB_FORCE) k_force_v2(Force2Args a) {
    __shared__ int s_q[QCAP][TPB_FORCE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[2] = {0.0, 0.0};
    const bool active = i < a.n && !xf_is_ghost(a.xf[i]);
    bool interior = *a.unwrapped == 0;
    if (active) interior = interior && is_interior(a.boxf, a.xf[i]);
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    if (active) {
        double fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0, vir = 0.0;
        if (warp_interior) force2_body<MULTI, false>(a, i, s_q, fx, fy, fz, pe, vir);
        else force2_body<MULTI, true>(a, i, s_q, fx, fy, fz, pe, vir);
        if (a.ax) {
            fx += a.ax[i];
            fy += a.ay[i];
            fz += a.az[i];
        }
        a.fx[i] = fx;
        a.fy[i] = fy;
        a.fz[i] = fz;
        red[0] = pe;
        red[1] = vir;
    }
    pisb_thermo *th = a.thermo;
    block_reduce_finalize<2, TPB_FORCE>(red, a.partials, a.ticket, [&](int qq, double s) {
        if (qq == 0) th->pe = s / 2.0;
        else th->virial_pair = s / 2.0;
    });
}

struct Build2Args {
    int n, npad, kcap;
    const double4 *xt;
    const float4 *xf;
    const int *cell_start;
    BoxDev box;
    BoxF boxf;
    Grid g;
    PairDev pair0;
    PairF pairf0;
    const PairDev *table;
    const PairF *tablef;
    int n_types;
    int *nbr;
    int *nnbr;
    int *flags;
};

template <bool MULTI, bool IMAGE>
__device__ __forceinline__ int build2_body(const Build2Args &a, int i) {
    int cnt = 0;
    const double4 xi = a.xt[i];
    const float4 xif = a.xf[i];
    const int ti = MULTI ? xf_type(xif) : 1;
    int c[3];
    cell_coords<true>(a.box, a.g, xi.x, xi.y, xi.z, c);
    int *const tile0 = a.nbr + (size_t)i * 4;
    const size_t tile_stride = (size_t)a.npad * 4;
    auto test = [&](int jj, const float4 &xjf) {
        const float r2f = r2_f32<IMAGE>(a.boxf, xif, xjf);
        PairF pf;
        int pidx = 0;
        if (MULTI) {
            const int tj = xf_type(xjf);
            pidx = (min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1);
            pf = a.tablef[pidx];
        } else {
            pf = a.pairf0;
        }
        bool in = r2f < pf.lo_list;
        if (!in && r2f <= pf.hi_list) {  // guard band: the exact FP64 reference predicate decides
            const double4 xj = ldg_d4(&a.xt[jj]);
            const double t_list = MULTI ? a.table[pidx].t_list : a.pair0.t_list;
            in = !(r2_exact_ortho(a.box, xi, xj) > t_list);
        }
        if (in && jj != i) {
            if (cnt < a.kcap) tile0[(size_t)(cnt >> 2) * tile_stride + (cnt & 3)] = jj;
            ++cnt;
        }
    };
    // candidates of a contiguous slot range [jb, je): 4 independent loads in flight, scalar tail
    auto scan = [&](int jb, int je) {
        int j = jb;
        for (; j + 4 <= je; j += 4) {
            float4 xjf[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xjf[u] = __ldg(&a.xf[j + u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) test(j + u, xjf[u]);
        }
        for (; j < je; ++j) test(j, __ldg(&a.xf[j]));
    };
    const int nx = a.g.n[0];
    for (int dz = a.g.lo[2]; dz <= a.g.hi[2]; ++dz) {
        int cz = c[2] + dz;
        if (cz < 0 || cz >= a.g.n[2]) {
            if (a.g.local[2]) continue;
            cz += cz < 0 ? a.g.n[2] : -a.g.n[2];
        }
        for (int dy = a.g.lo[1]; dy <= a.g.hi[1]; ++dy) {
            int cy = c[1] + dy;
            if (cy < 0 || cy >= a.g.n[1]) {
                if (a.g.local[1]) continue;
                cy += cy < 0 ? a.g.n[1] : -a.g.n[1];
            }
            const int rb = (cz * a.g.n[1] + cy) * nx;
            // the x-neighbour cells of one row are consecutive slots: same visiting order as
            // dx = lo..hi with periodic wrap (wrapped-low part, main part, wrapped-high part)
            int xlo = c[0] + a.g.lo[0];
            const int xhi = c[0] + a.g.hi[0];
            if (xlo < 0) {
                if (!a.g.local[0]) scan(__ldg(&a.cell_start[rb + xlo + nx]), __ldg(&a.cell_start[rb + nx]));
                xlo = 0;
            }
            const int xhi_main = min(xhi, nx - 1);
            scan(__ldg(&a.cell_start[rb + xlo]), __ldg(&a.cell_start[rb + xhi_main + 1]));
            if (xhi >= nx && !a.g.local[0]) scan(__ldg(&a.cell_start[rb]), __ldg(&a.cell_start[rb + xhi - nx + 1]));
        }
    }
    nbr_close_row(a.nbr, i, a.npad, cnt, a.kcap);
    a.nnbr[i] = cnt < a.kcap ? cnt : a.kcap;
    return cnt;
}

template <bool MULTI>
__global__ void __launch_bounds__(TPB_FORCE) k_build_list_v2(Build2Args a) {
    // one thread per atom on a full grid: a capped grid-stride version was measured 20 % slower (tail + locality),
    // so the no-rebuild early exit of this one kernel costs ~20 us per step
    if (a.flags[FLAG_REBUILD] == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n && xf_is_ghost(a.xf[i])) a.nnbr[i] = 0;
    const bool active = i < a.n && !xf_is_ghost(a.xf[i]);
    bool interior = true;
    if (active) interior = is_interior(a.boxf, a.xf[i]);
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    int cnt = 0;
    if (active) cnt = warp_interior ? build2_body<MULTI, false>(a, i) : build2_body<MULTI, true>(a, i);
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&a.flags[FLAG_MAXNBR], m);
}


// ------------------------------------------------------------------------------------------------
// k_build_list_v3: the v2 build with the inner loop rebuilt around Blackwell's packed FP32 pipe.
// ncu on v2 (profiles/r01_prof_build_v2b.*): 25 warp-instructions per candidate, issue-bound, the append block
// (address arithmetic + predicated store) issued for almost every candidate because SOME lane accepts.  v3, for
// interior warps of a single-type system:
//   * candidates come as PAIRS: slots 2q, 2q+1 share one 32-byte record {x0,x1,y0,y1,z0,z1,w0,w1} (xp, written by
//     k_copy_back at rebuild time), fetched with one 256-bit load; the three differences and the squared distance of
//     BOTH candidates are 6 packed instructions (FADD2 x3, FMUL2, FFMA2 x2) instead of 12 scalar ones;
//   * acceptance is recorded as bits of a per-thread 32-candidate mask (FSETP + predicated LOP3), a second mask marks
//     r2 < lo: their difference is the FP32 guard band, resolved by the exact FP64 reference predicate (rare);
//   * the list is appended once per 32 candidates by walking the set bits, so the store path is issued ~max popc
//     times per chunk instead of once per candidate; range edges are a mask, the self pair is a split range.
// The FP32 operations are the same IEEE operations in the same order as r2_f32<false>, so the decision is
// bit-for-bit the one v2 makes (and therefore the reference's `rij.norm() > rcut`).
// Boundary warps (image needed), multi-type tables and brick-local grids take the v2 body.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_neg(f32x2_t a) { return a ^ 0x8000000080000000ull; }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
struct PairRec {
    f32x2_t x, y, z, w;
};
__device__ __forceinline__ PairRec ldg_pair(const float *xp, int q) {
    PairRec r;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w) : "l"(xp + (size_t)q * 8));
    return r;
}


// The exact reference predicate for a guard-band candidate (rare).  Out of line, by-value arguments only (taking the
// address of a kernel parameter would force a local copy of the whole argument struct).  v2_possible() => fully periodic.
__device__ __noinline__ bool build3_exact_out(double h0, double h1, double h2, double i0, double i1, double i2,
                                              const double4 *__restrict__ xt, double xix, double xiy, double xiz, int jj,
                                              double t_list) {
    const double4 xj = ldg_d4(&xt[jj]);
    double sx = __dmul_rn(i0, __dsub_rn(xj.x, xix)), sy = __dmul_rn(i1, __dsub_rn(xj.y, xiy)), sz = __dmul_rn(i2, __dsub_rn(xj.z, xiz));
    sx = __dsub_rn(sx, round(sx));
    sy = __dsub_rn(sy, round(sy));
    sz = __dsub_rn(sz, round(sz));
    return norm2(__dmul_rn(h0, sx), __dmul_rn(h1, sy), __dmul_rn(h2, sz)) > t_list;
}

// B3_PAIRS: pair records in flight per thread (2 candidates each).  Measured at 4M atoms in the melt, ms per build
// (profiles/r01_build_occupancy.jsonl): 64 registers / 4 records 2.48, 48 registers / 4 records 2.31, 48 registers /
// 2 records 2.30 (the default), 40 registers 2.34, 8 records 2.74; loading the next stencil row's cell_start words while
// the current row is scanned changed nothing.
// MULTI: per-pair cutoffs.  The block keeps the FP32 band {lo, hi} of every type pair in shared memory as a symmetric
// (n_types + 2)^2 table whose row / column 0 and n_types + 1 hold {-1, -1} (nothing is accepted: padding slots carry type 0,
// stale padding anything), so a candidate costs two LDS.64 on top of the single-type loop; a missing pair is {-1, -1} too.
constexpr int B3_MAX_TYPES = 6;
template <bool IMAGE, bool MULTI, int B3_PAIRS = 2>
__device__ __forceinline__ int build3_body(const Build2Args &a, const float *__restrict__ xp, int i, const float2 *s_band) {
    int cnt = 0;
    const double4 xi = a.xt[i];
    const float4 xif = a.xf[i];
    const int tdim = a.n_types + 2;
    const float2 *band_row = MULTI ? s_band + xf_type(xif) * tdim : nullptr;
    int c[3];
    cell_coords<true>(a.box, a.g, xi.x, xi.y, xi.z, c);
    int *wp = a.nbr + (size_t)i * 4;  // slot of list entry number cnt
    const ptrdiff_t tile_step = (ptrdiff_t)a.npad * 4 - 3;
    const float lo = a.pairf0.lo_list, hi = a.pairf0.hi_list;
    const float Lx = a.boxf.L[0], Ly = a.boxf.L[1], Lz = a.boxf.L[2];
    const float iLx = a.boxf.invL[0], iLy = a.boxf.invL[1], iLz = a.boxf.invL[2];
    const f32x2_t xi2 = f2_pack(xif.x, xif.x), yi2 = f2_pack(xif.y, xif.y), zi2 = f2_pack(xif.z, xif.z);
    const f32x2_t magic = f2_pack(12582912.0f, 12582912.0f);
    // d - L * rint(d / L) for both candidates, the same FP32 operations as r2_f32<true> (rint by magic constant)
    auto image = [&](f32x2_t d, float L, float iL) {
        const f32x2_t t = f2_mul(d, f2_pack(iL, iL));
        const f32x2_t r = f2_sub(f2_add(t, magic), magic);
        return f2_fma(f2_neg(r), f2_pack(L, L), d);
    };
    // candidates of the slot range [jb, je)
    auto scan = [&](int jb, int je) {
        const int q_end = (je + 1) >> 1;
        for (int q0 = jb >> 1; q0 < q_end; q0 += 16) {
            unsigned m_hi = 0u, m_lo = 0u;  // bit b <-> candidate slot 2*q0 + b
#pragma unroll 1
            for (int s = 0; s < 16 && q0 + s < q_end; s += B3_PAIRS) {
                PairRec r[B3_PAIRS];
#pragma unroll
                for (int u = 0; u < B3_PAIRS; ++u) r[u] = ldg_pair(xp, q0 + s + u);  // array padded: reads past q_end stay in bounds
                unsigned b_hi = 0u, b_lo = 0u;
#pragma unroll
                for (int u = 0; u < B3_PAIRS; ++u) {
                    f32x2_t dx = f2_sub(r[u].x, xi2), dy = f2_sub(r[u].y, yi2), dz = f2_sub(r[u].z, zi2);
                    if (IMAGE) {
                        dx = image(dx, Lx, iLx);
                        dy = image(dy, Ly, iLy);
                        dz = image(dz, Lz, iLz);
                    }
                    const f32x2_t r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
                    float ra, rb;
                    f2_unpack(r2, ra, rb);
                    float lo_a = lo, hi_a = hi, lo_b = lo, hi_b = hi;
                    if (MULTI) {
                        float wa, wb;
                        f2_unpack(r[u].w, wa, wb);
                        const float2 ba = band_row[min((unsigned)xf_type(make_float4(0.f, 0.f, 0.f, wa)), (unsigned)(tdim - 1))];
                        const float2 bb = band_row[min((unsigned)xf_type(make_float4(0.f, 0.f, 0.f, wb)), (unsigned)(tdim - 1))];
                        lo_a = ba.x, hi_a = ba.y, lo_b = bb.x, hi_b = bb.y;
                    }
                    if (ra <= hi_a) b_hi |= 1u << (2 * u);
                    if (rb <= hi_b) b_hi |= 2u << (2 * u);
                    if (ra < lo_a) b_lo |= 1u << (2 * u);
                    if (rb < lo_b) b_lo |= 2u << (2 * u);
                }
                m_hi |= b_hi << (2 * s);
                m_lo |= b_lo << (2 * s);
            }
            // range edges (and whatever lies in the padding / beyond q_end)
            const int first = jb - 2 * q0, last = je - 2 * q0;  // valid bits: [first, last)
            unsigned valid = last >= 32 ? 0xffffffffu : ((1u << last) - 1u);
            if (first > 0) valid &= ~((1u << first) - 1u);
            const unsigned self = (unsigned)(i - 2 * q0);  // the atom itself, when this chunk holds it, is no neighbour
            if (self < 32u) valid &= ~(1u << self);
            m_hi &= valid;
            // FP32 guard band: the exact FP64 reference predicate decides
            unsigned band = m_hi & ~m_lo;
            while (band) {
                const int b = __ffs(band) - 1;
                band &= band - 1u;
                double t_list = a.pair0.t_list;
                if (MULTI) {
                    const int ti = xf_type(xif), tj = xf_type(a.xf[2 * q0 + b]);
                    t_list = a.table[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)].t_list;
                }
                if (build3_exact_out(a.box.h[0], a.box.h[4], a.box.h[8], a.box.hinv[0], a.box.hinv[4], a.box.hinv[8], a.xt, xi.x,
                                     xi.y, xi.z, 2 * q0 + b, t_list))
                    m_hi &= ~(1u << b);
            }
            const int j0 = 2 * q0;
            while (m_hi) {
                const int b = __ffs(m_hi) - 1;
                m_hi &= m_hi - 1u;
                if (cnt < a.kcap) *wp = j0 + b;
                ++cnt;
                wp += (cnt & 3) ? (ptrdiff_t)1 : tile_step;
            }
        }
    };
    const int nx = a.g.n[0];
    for (int dz = a.g.lo[2]; dz <= a.g.hi[2]; ++dz) {
        int cz = c[2] + dz;
        if (cz < 0 || cz >= a.g.n[2]) {
            // interior atom: a wrapped cell lies beyond the cutoff (margin > rc + skin); brick-local dims never wrap
            if (!IMAGE || a.g.local[2]) continue;
            cz += cz < 0 ? a.g.n[2] : -a.g.n[2];
        }
        for (int dy = a.g.lo[1]; dy <= a.g.hi[1]; ++dy) {
            int cy = c[1] + dy;
            if (cy < 0 || cy >= a.g.n[1]) {
                if (!IMAGE || a.g.local[1]) continue;
                cy += cy < 0 ? a.g.n[1] : -a.g.n[1];
            }
            const int rb = (cz * a.g.n[1] + cy) * nx;
            const int xlo = c[0] + a.g.lo[0], xhi = c[0] + a.g.hi[0];
            if (IMAGE && xlo < 0 && !a.g.local[0]) scan(__ldg(&a.cell_start[rb + xlo + nx]), __ldg(&a.cell_start[rb + nx]));
            const int jb = __ldg(&a.cell_start[rb + max(xlo, 0)]);
            const int je = __ldg(&a.cell_start[rb + min(xhi, nx - 1) + 1]);
            scan(jb, je);  // (i itself is masked out inside: splitting the range at i cost the centre row a second chunk)
            if (IMAGE && xhi >= nx && !a.g.local[0]) scan(__ldg(&a.cell_start[rb]), __ldg(&a.cell_start[rb + xhi - nx + 1]));
        }
    }
    nbr_close_row(a.nbr, i, a.npad, cnt, a.kcap);
    a.nnbr[i] = cnt < a.kcap ? cnt : a.kcap;
    return cnt;
}

// 10 resident blocks per SM (48 registers, 40 warps): the kernel waits on its record loads (ncu: 44 % of the stall cycles
// long-scoreboard at 6.6 warps per scheduler), so occupancy beats the handful of spills it costs.
template <bool MULTI>
__global__ void __launch_bounds__(TPB_FORCE, 10) k_build_list_v3(Build2Args a, const float *__restrict__ xp, int fast_ok) {
    if (a.flags[FLAG_REBUILD] == 0) return;
    __shared__ float2 s_band[MULTI ? (B3_MAX_TYPES + 2) * (B3_MAX_TYPES + 2) : 1];
    if (MULTI && fast_ok) {
        const int tdim = a.n_types + 2;
        for (int k = threadIdx.x; k < tdim * tdim; k += blockDim.x) {
            const int ti = k / tdim, tj = k % tdim;
            float2 v = make_float2(-1.f, -1.f);
            if (ti >= 1 && ti <= a.n_types && tj >= 1 && tj <= a.n_types) {
                const PairF pf = a.tablef[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)];
                v = make_float2(pf.lo_list, pf.hi_list);
            }
            s_band[k] = v;
        }
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n && xf_is_ghost(a.xf[i])) a.nnbr[i] = 0;
    const bool active = i < a.n && !xf_is_ghost(a.xf[i]);
    bool interior = true;
    if (active) interior = is_interior(a.boxf, a.xf[i]);
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    int cnt = 0;
    if (active) {
        if (fast_ok) cnt = warp_interior ? build3_body<false, MULTI>(a, xp, i, s_band) : build3_body<true, MULTI>(a, xp, i, s_band);
        else cnt = warp_interior ? build2_body<MULTI, false>(a, i) : build2_body<MULTI, true>(a, i);
    }
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&a.flags[FLAG_MAXNBR], m);
}

// ------------------------------------------------------------------------------------------------
// pisb_list_stats: what the CURRENT list holds, counted on the device -- listed pairs (sum of the per-atom row lengths)
// and pairs inside the cutoff right now (the reference predicate).  bench.py's roofline needs K and K_in.
// ------------------------------------------------------------------------------------------------
struct ListStatsArgs {
    int n, npad;
    const double4 *xt;
    const int *nbr, *nnbr;
    BoxDev box;
    PairDev pair0;
    const PairDev *table;
    int n_types;
    unsigned long long *out;  // [0] listed, [1] in range
};

template <bool ORTHO>
__global__ void __launch_bounds__(TPB) k_list_stats(ListStatsArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long listed = 0, inr = 0;
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double4 xi = a.xt[i];
        const int ti = type_of(xi.w);
        const int nn = a.nnbr[i];
        listed = (unsigned long long)nn;
        for (int k = 0; k < nn; ++k) {
            const double4 xj = ldg_d4(&a.xt[a.nbr[nbr_at(k, i, a.npad)]]);
            double dx = __dsub_rn(xj.x, xi.x), dy = __dsub_rn(xj.y, xi.y), dz = __dsub_rn(xj.z, xi.z);
            min_image<ORTHO>(a.box, dx, dy, dz);
            const int tj = type_of(xj.w);
            const PairDev pd = a.n_types > 1 ? a.table[(min(ti, tj) - 1) * a.n_types + (max(ti, tj) - 1)] : a.pair0;
            if (pd.present && !(norm2(dx, dy, dz) > pd.t_rc)) ++inr;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        listed += __shfl_down_sync(0xffffffffu, listed, o);
        inr += __shfl_down_sync(0xffffffffu, inr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.out[0], listed);
        atomicAdd(&a.out[1], inr);
    }
}

}  // namespace pisb
