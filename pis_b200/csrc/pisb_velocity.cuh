// pisb_velocity.cuh -- `velocity all create T seed` on the device.
//
// Replaces Atoms::start_velocities (src/atoms/velocities.rs:10-15) for a state that already lives in HBM:
//   initialise_velocities  :17-33   v_i ~ Normal(0, sqrt(kB T / m_i)) per component
//   remove_drift           :35-50   v -= (sum m v) / (sum m)
//   rescale_to_temperature :52-59   v *= sqrt(T / T_now),  T_now = 2 KE / (3 N kB)   (properties.rs:17-30,41-43)
// The reference's stream (rand SmallRng + rand_distr Normal) is third-party and unpinned, so the Gaussian
// numbers come from this repository's generator: splitmix64 keyed by (seed, GLOBAL atom id, component) + Box-Muller --
// the one the C++ host (pis_host.cpp) and pis_b200/decomposition.py define.  Counter-based on purpose: a brick of a
// multi-GPU run generates exactly the values its atoms would get on one GPU, whatever the slot order.
// Three launches, each a deterministic three-level reduction (pisb_device.cuh); T = 0 gives NaN velocities like the
// reference (0/0 in the rescale).
#pragma once
#include "pisb_kernels.cuh"

namespace pisb {

__device__ __forceinline__ unsigned long long splitmix64_dev(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct VelInitArgs {
    int n;
    const double4 *xt;  // type + ghost flag of every slot
    const int *id;      // global atom id of every slot
    double *vx, *vy, *vz;
    const double *mass;
    double kt;                // kB * T
    unsigned long long base;  // splitmix64(seed * 0xD1342543DE82EF95 + 12345)
    double temperature;
    double *sums;  // [0] sum m, [1..3] sum m v, [4] owned atoms, [5] KE after drift removal
    double *partials;
    unsigned int *ticket;
};

// velocities.rs:17-33 + the sums remove_drift needs
__global__ void __launch_bounds__(TPB) k_vel_create(VelInitArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double m = a.mass[type_of(a.xt[i].w) - 1];
        const double sigma = sqrt(a.kt / m);
        const unsigned long long key = (unsigned long long)(unsigned int)a.id[i] * 6ULL + a.base;
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned long long k = splitmix64_dev(key + 2 * c), k2 = splitmix64_dev(key + 2 * c + 1);
            const double u1 = ((double)(k >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
            const double u2 = (double)(k2 >> 11) * (1.0 / 9007199254740992.0);         // [0, 1)
            v[c] = sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2) * sigma;
        }
        a.vx[i] = v[0];
        a.vy[i] = v[1];
        a.vz[i] = v[2];
        red[0] = m;
        red[1] = v[0] * m;
        red[2] = v[1] * m;
        red[3] = v[2] * m;
        red[4] = 1.0;
    }
    double *sums = a.sums;
    block_reduce_finalize<5, TPB>(red, a.partials, a.ticket, [&](int q, double s) { sums[q] = s; });
}

// velocities.rs:35-50, + the kinetic energy rescale_to_temperature starts from (properties.rs:17-24)
__global__ void __launch_bounds__(TPB) k_vel_remove_drift(VelInitArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[1] = {0.0};
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double tm = a.sums[0];
        const double vx = a.vx[i] - (a.sums[1] / tm) * 1.0, vy = a.vy[i] - (a.sums[2] / tm) * 1.0, vz = a.vz[i] - (a.sums[3] / tm) * 1.0;
        a.vx[i] = vx;
        a.vy[i] = vy;
        a.vz[i] = vz;
        const double m = a.mass[type_of(a.xt[i].w) - 1];
        red[0] = __dmul_rn(__dmul_rn(0.5, m), norm2(vx, vy, vz));
    }
    double *sums = a.sums;
    block_reduce_finalize<1, TPB>(red, a.partials, a.ticket, [&](int, double s) { sums[5] = s; });
}

// velocities.rs:52-59
__global__ void __launch_bounds__(TPB) k_vel_rescale(VelInitArgs a, double kb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double t_now = (2.0 * a.sums[5]) / ((3.0 * a.sums[4]) * kb);
        const double lambda = sqrt(a.temperature / t_now);
        a.vx[i] *= lambda;
        a.vy[i] *= lambda;
        a.vz[i] *= lambda;
    }
}

}  // namespace pisb
