// pisb_multi.cuh -- device kernels of the multi-GPU path (spatial decomposition, one process per GPU).
//
// NEW functionality (the reference is single-process).  The periodic box is cut into
// P[0] x P[1] x P[2] bricks (P_d in {1, 2}: 1, 2x1x1, 2x2x1, 2x2x2 on one NVSwitch node, where every
// GPU is one hop from every other).  A rank OWNS the atoms whose wrapped position lies in its brick
// and keeps GHOST copies of foreign atoms within rc + skin of the brick.  Positions always stay
// GLOBAL wrapped coordinates and the pair arithmetic uses the global box, so every pair term is
// bit-identical to the single-GPU (and reference) value; with a FULL neighbour list each rank
// computes complete forces for its owned atoms and no reverse (force) exchange exists.
//   every step      : ghost positions, gathered and STORED straight into the peers' receive buffers over NVLink
//                     (k_halo_push on CUDA-IPC mapped peer memory, sequence-number signals, k_halo_pull scatters
//                     into the cell-sorted slots and max-reduces the skin-trigger flags); grouped
//                     ncclSend/ncclRecv (k_halo_pack / k_halo_unpack) is the fallback when IPC mapping fails
//   rebuild steps   : atom migration, ghost selection, then the ordinary bin/sort/build chain
//   end of a batch  : one sum-allreduce of all per-step thermo records
#pragma once
#include "pisb_kernels.cuh"

namespace pisb {

struct Decomp {
    int P[3];        // bricks per dimension (1 or 2)
    int b[3];        // this rank's brick coordinates
    int rank, nranks;
    double lo[3], hi[3];  // brick bounds, global coordinates
    double gw;            // ghost width (rc + skin, padded)
};

constexpr int MIG_REC = 11;    // x y z w vx vy vz fx fy fz id
constexpr int GHOST_REC = 5;   // x y z w id

__device__ __forceinline__ int brick_rank(const Decomp &dc, const int *bb) { return (bb[2] * dc.P[1] + bb[1]) * dc.P[0] + bb[0]; }

__device__ __forceinline__ int dest_rank(const BoxDev &box, const Decomp &dc, const double4 &x) {
    const double r[3] = {x.x, x.y, x.z};
    int bb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double s = r[d] * box.hinv[4 * d];
        s -= floor(s);
        int bd = (int)floor(s * (double)dc.P[d]);
        bb[d] = min(max(bd, 0), dc.P[d] - 1);
    }
    return brick_rank(dc, bb);
}

// ---- migration ---------------------------------------------------------------------------------
// Only atoms that LEAVE are touched: they are packed for their new owner and their slot is marked
// dead; stale ghosts are marked dead too.  Arrivals and fresh ghosts are appended behind the live
// slots and the ordinary cell sort then compacts (dead slots land in the sentinel bucket).
__global__ void __launch_bounds__(TPB) k_mig_count(int n, const double4 *__restrict__ xt, BoxDev box, Decomp dc,
                                                   int *__restrict__ dest, int *__restrict__ pig, int *__restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 x = xt[i];
    int d = -1;
    if (!is_ghost(x.w)) {
        d = dest_rank(box, dc, x);
        if (d == dc.rank) d = -1;
    }
    dest[i] = d;
    if (d >= 0) pig[i] = atomicAdd(&cnt[d], 1);  // leavers are rare: no contention
}

struct MigArgs {
    int n;
    double4 *xt;
    const double *vx, *vy, *vz, *fx, *fy, *fz;
    const int *id;
    const int *dest, *pig, *off;
    double *buf;
};

__global__ void __launch_bounds__(TPB) k_mig_pack(MigArgs a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int d = a.dest[i];
    double4 x = a.xt[i];
    if (d < 0) {
        if (is_ghost(x.w) && !is_dead(x.w)) {  // stale ghost: re-selected below
            x.w = mark_dead(x.w);
            a.xt[i] = x;
        }
        return;
    }
    double *r = a.buf + (size_t)(a.off[d] + a.pig[i]) * MIG_REC;
    r[0] = x.x;
    r[1] = x.y;
    r[2] = x.z;
    r[3] = x.w;
    r[4] = a.vx[i];
    r[5] = a.vy[i];
    r[6] = a.vz[i];
    r[7] = a.fx[i];
    r[8] = a.fy[i];
    r[9] = a.fz[i];
    r[10] = __longlong_as_double((long long)a.id[i]);
    x.w = mark_dead(x.w);
    a.xt[i] = x;
}

struct MigUnpackArgs {
    int count, dst;
    const double *buf;
    double4 *xt;
    double *vx, *vy, *vz, *fx, *fy, *fz;
    int *id;
};

__global__ void __launch_bounds__(TPB) k_mig_unpack(MigUnpackArgs a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.count) return;
    const double *r = a.buf + (size_t)k * MIG_REC;
    const int s = a.dst + k;
    a.xt[s] = make_double4(r[0], r[1], r[2], r[3]);
    a.vx[s] = r[4];
    a.vy[s] = r[5];
    a.vz[s] = r[6];
    a.fx[s] = r[7];
    a.fy[s] = r[8];
    a.fz[s] = r[9];
    a.id[s] = (int)__double_as_longlong(r[10]);
}

// ---- ghost selection ---------------------------------------------------------------------------
// An owned atom is sent to the rank reached by flipping the brick coordinate in every dimension of a
// non-empty subset S of the decomposed dimensions iff it is within gw of a face in every d in S.
// FILL == false: count per destination; FILL == true: write the send list.
template <bool FILL>
__global__ void __launch_bounds__(TPB) k_ghost_select(int n, const double4 *__restrict__ xt, Decomp dc,
                                                      int *__restrict__ cnt, const int *__restrict__ off,
                                                      int *__restrict__ send_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 x = xt[i];
    if (is_ghost(x.w)) return;
    const double r[3] = {x.x, x.y, x.z};
    int near = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        if (dc.P[d] > 1 && (r[d] - dc.lo[d] < dc.gw || dc.hi[d] - r[d] < dc.gw)) near |= 1 << d;
    if (!near) return;
    for (int sset = 1; sset < 8; ++sset) {
        if ((sset & near) != sset) continue;
        int bb[3] = {dc.b[0], dc.b[1], dc.b[2]};
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (sset & (1 << d)) bb[d] ^= 1;
        const int dst = brick_rank(dc, bb);
        const int p = atomicAdd(&cnt[dst], 1);
        if (FILL) send_idx[off[dst] + p] = i;
    }
}

__global__ void __launch_bounds__(TPB) k_ghost_pack_full(int count, const int *__restrict__ send_idx,
                                                         const double4 *__restrict__ xt, const int *__restrict__ id,
                                                         double *__restrict__ buf) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int s = send_idx[k];
    const double4 x = xt[s];
    double *r = buf + (size_t)k * GHOST_REC;
    r[0] = x.x;
    r[1] = x.y;
    r[2] = x.z;
    r[3] = type_ghost_as_double(type_of(x.w), true);
    r[4] = __longlong_as_double((long long)id[s]);
}

__global__ void __launch_bounds__(TPB) k_ghost_append(int count, int dst, const double *__restrict__ buf,
                                                      double4 *__restrict__ xt, double *vx, double *vy, double *vz,
                                                      double *fx, double *fy, double *fz, int *__restrict__ id) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const double *r = buf + (size_t)k * GHOST_REC;
    const int s = dst + k;
    xt[s] = make_double4(r[0], r[1], r[2], r[3]);
    vx[s] = vy[s] = vz[s] = 0.0;
    fx[s] = fy[s] = fz[s] = 0.0;
    id[s] = (int)__double_as_longlong(r[4]);
}

// after the cell sort: newslot[old] = new
__global__ void __launch_bounds__(TPB) k_inverse_perm(int n, const int *__restrict__ order, int *__restrict__ newslot) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) newslot[order[p]] = p;
}

__global__ void __launch_bounds__(TPB) k_remap(int count, int *__restrict__ idx, const int *__restrict__ newslot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) idx[k] = newslot[idx[k]];
}

__global__ void __launch_bounds__(TPB) k_ghost_slots(int count, int first_old, const int *__restrict__ newslot,
                                                     int *__restrict__ ghost_slot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) ghost_slot[k] = newslot[first_old + k];
}

// ---- per-step halo ------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_halo_pack(int count, const int *__restrict__ send_idx,
                                                   const double4 *__restrict__ xt, double4 *__restrict__ buf) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) buf[k] = xt[send_idx[k]];
}

// flag_in[r] = skin-trigger flag received from rank r (exchanged in the same NCCL group as the ghosts):
// the global decision max_r(flag) is formed here, so no separate all-reduce is needed (with P_d <= 2
// every rank is a peer of every other).
__global__ void __launch_bounds__(TPB) k_halo_unpack(int count, const int *__restrict__ ghost_slot,
                                                     const double4 *__restrict__ buf, double4 *__restrict__ xt,
                                                     float4 *__restrict__ xf, BoxDev box, const int *__restrict__ flag_in,
                                                     int nranks, int me, int *__restrict__ flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && flag_in) {
        int f = flags[FLAG_REBUILD];
        for (int r = 0; r < nranks; ++r)
            if (r != me) f = max(f, flag_in[r]);
        flags[FLAG_REBUILD] = f;
    }
    if (k >= count) return;
    const int s = ghost_slot[k];
    double4 x = buf[k];
    x.w = type_ghost_as_double(type_of(x.w), true);
    xt[s] = x;
    xf[s] = make_xf(box, x);
}

// ---- per-step halo over peer memory (NVLink stores; no NCCL in the step loop) ---------------------------
// Every rank maps every peer's receive buffer and signal words (CUDA IPC).  One exchange, sequence number seq:
//   k_halo_push   : gathers this rank's ghost-source positions and STORES them straight into the peers' receive
//                   buffers (slot [seq&1]: double-buffered, a rank can be at most one exchange ahead of a peer
//                   because its next force step needs that peer's data);
//   k_halo_signal : after the stores, publishes (seq << 1 | skin-trigger flag) into each peer's signal word;
//   k_halo_pull   : spins until every peer's signal for seq has arrived, scatters the received records into the
//                   cell-sorted ghost slots, and max-reduces the flags (the global rebuild decision).
constexpr int P2P_MAX_RANKS = 8;

struct P2PArgs {
    int nranks, me, send_total, n_ghost;
    int half;                        // records per parity half of every receive buffer (same on all ranks)
    int src_off[P2P_MAX_RANKS];      // first send entry for peer r
    int cnt[P2P_MAX_RANKS];          // entries for peer r
    int dst_off[P2P_MAX_RANKS];      // where this rank's block starts inside peer r's receive layout
    double4 *peer_recv[P2P_MAX_RANKS];
    unsigned long long *peer_sig[P2P_MAX_RANKS];
    const double4 *my_recv;
    const unsigned long long *my_sig;  // [2][P2P_MAX_RANKS]
};

__global__ void __launch_bounds__(TPB) k_halo_push(P2PArgs a, unsigned long long seq, const int *__restrict__ send_idx,
                                                   const double4 *__restrict__ xt) {
    const int par = (int)(seq & 1ull);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.send_total; k += gridDim.x * blockDim.x) {
        int r = 0;
#pragma unroll
        for (int q = 0; q < P2P_MAX_RANKS; ++q)
            if (q < a.nranks && a.cnt[q] > 0 && k >= a.src_off[q]) r = q;  // send entries are grouped by destination rank
        a.peer_recv[r][(size_t)par * a.half + a.dst_off[r] + (k - a.src_off[r])] = xt[send_idx[k]];
    }
}

__global__ void k_halo_signal(P2PArgs a, unsigned long long seq, const int *__restrict__ flags) {
    const int r = threadIdx.x;
    if (r >= a.nranks || r == a.me) return;
    __threadfence_system();
    const unsigned long long v = (seq << 1) | (unsigned long long)(flags[FLAG_REBUILD] ? 1 : 0);
    atomicExch_system(&a.peer_sig[r][(seq & 1ull) * P2P_MAX_RANKS + a.me], v);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(TPB) k_halo_pull(P2PArgs a, unsigned long long seq, const int *__restrict__ ghost_slot,
                                                   double4 *__restrict__ xt, float4 *__restrict__ xf, BoxDev box,
                                                   int *__restrict__ flags, long long timeout_cycles) {
    __shared__ int s_flag, s_timeout;
    const int par = (int)(seq & 1ull);
    if (threadIdx.x == 0) {
        int f = 0, to = 0;
        const long long t0 = clock64();
        for (int r = 0; r < a.nranks; ++r) {
            if (r == a.me) continue;
            unsigned long long v;
            while (((v = ld_acquire_sys(&a.my_sig[par * P2P_MAX_RANKS + r])) >> 1) < seq) {
                if (clock64() - t0 > timeout_cycles) {
                    to = 1;
                    break;
                }
                __nanosleep(100);
            }
            f |= (int)(v & 1ull);
        }
        s_flag = f;
        s_timeout = to;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (s_flag) flags[FLAG_REBUILD] = 1;
        if (s_timeout) flags[FLAG_COMM_TIMEOUT] = 1;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.n_ghost; k += gridDim.x * blockDim.x) {
        const int s = ghost_slot[k];
        const double2 *rp = reinterpret_cast<const double2 *>(&a.my_recv[(size_t)par * a.half + k]);
        const double2 lo = __ldcg(rp), hi = __ldcg(rp + 1);  // L2 only: the data was written by a peer over NVLink
        double4 x = make_double4(lo.x, lo.y, hi.x, hi.y);
        x.w = type_ghost_as_double(type_of(x.w), true);
        xt[s] = x;
        xf[s] = make_xf(box, x);
    }
}

// ---- the per-step decision, published without a copy engine -------------------------------------
// One warp copies flags[] into page-locked HOST memory (directly addressable under UVA) and then a sequence number the
// host spins on.  A cudaMemcpyAsync of the same 32 bytes shares the device-to-host copy engine with an asynchronous dump
// frame (pisb_download_owned_begin, 112 MB per rank) and waits behind it: measured at 8 GPUs, the 10-step batch that runs
// under a frame took 28 instead of 19 ms.  With FUSED the stream-ordered copy FLAG_DECISION <- FLAG_REBUILD is made here too.
__global__ void k_publish_flags(int *__restrict__ flags, volatile int *__restrict__ host, int seq, int set_decision) {
    const int t = threadIdx.x;
    if (set_decision && t == 0) flags[FLAG_DECISION] = flags[FLAG_REBUILD];
    __syncwarp();
    if (t < FLAG_COUNT) host[t] = flags[t];
    __threadfence_system();
    __syncwarp();
    if (t == 0) host[FLAG_COUNT] = seq;
}
// the same for any short word array (the rebuild's count tables): host[0..n) = src[0..n), then host[n] = seq; one warp
__global__ void k_publish_words(const int *__restrict__ src, volatile int *__restrict__ host, int n, int seq) {
    for (int k = threadIdx.x; k < n; k += 32) host[k] = src[k];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) host[n] = seq;
}

// ---- owned-atom download ----------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_store_owned(int n, const double4 *__restrict__ xt, const double *vx,
                                                     const double *vy, const double *vz, const double *fx,
                                                     const double *fy, const double *fz, const int *__restrict__ id,
                                                     int *__restrict__ counter, double *pos, double *vel, double *frc,
                                                     int *gid) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool own = s < n && !is_ghost(xt[s].w);
    // warp-aggregated slot allocation (ballot/popc): one atomic per warp instead of one per atom
    const unsigned m = __ballot_sync(0xffffffffu, own);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!own) return;
    const size_t o = (size_t)(base + __popc(m & ((1u << lane) - 1u)));
    const double4 x = xt[s];
    if (pos) {
        pos[3 * o] = x.x;
        pos[3 * o + 1] = x.y;
        pos[3 * o + 2] = x.z;
    }
    if (vel) {
        vel[3 * o] = vx[s];
        vel[3 * o + 1] = vy[s];
        vel[3 * o + 2] = vz[s];
    }
    if (frc) {
        frc[3 * o] = fx[s];
        frc[3 * o + 1] = fy[s];
        frc[3 * o + 2] = fz[s];
    }
    gid[o] = id[s];
}

}  // namespace pisb
