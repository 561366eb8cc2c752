// pisb_pairlist.cuh -- TWO ATOMS PER THREAD: the Verlet list and the force / step kernels of large single-type systems.
//
// Why.  ncu on the thread-per-atom force loop (profiles/r01_force_vv_full.*): FP64 pipe 66-74 %, and the L1TEX LSU data
// pipe at 90 % -- one data-pipe wavefront per distinct 128-byte line a warp's gather touches, ~25 per 256-bit gather
// instruction, because the k-th neighbours of 32 different atoms share no lines.  Trimming the FP64 work (the lean loop,
// pisb_kernels.cuh) does not move that second limit; the number of GATHERS has to come down.  A half list would halve
// them but needs F_j atomics (REDG f64: 1.29 cycles per lane spread, B300_MICROARCH.md -> 81 REDs per atom ~ 1.4 ms at 4M
// atoms, slower than the whole kernel) or a reverse pass that gathers just as much.  Instead:
//
//   thread p owns the atoms in slots 2p and 2p+1 -- consecutive slots of the cell-sorted order, i.e. the same or the
//   next cell, ~3.5 A apart, whose neighbour spheres (radius rc + skin = 9.5 A) overlap by ~73 %.  Its list has three
//   sections:   BOTH    slots that are neighbours of both atoms   (~63 entries)   one gather, two pair evaluations
//               ONLY_A  neighbours of 2p only                     (~23)          one gather, one evaluation
//               ONLY_B  neighbours of 2p+1 only                   (~23)
//   so the two atoms cost 63 + 23 + 23 = 109 gathers and index words instead of 2 x 86 = 172 (-37 %), with EXACTLY the
//   same pair evaluations as two per-atom rows (no cluster-pair inflation: the FP64 pipe is the other limit), the same
//   in/out decisions, and the same per-atom neighbour sets (pisb_neighbours reassembles BOTH + ONLY_x; the parity tests
//   compare them with the oracle's rows as before).
//
// Layout: three arrays shaped like the per-atom list, [kcap/4][npp][4] with npp = padded number of pair threads, carved out
// of the same allocation; counts[p] = {n_both, n_only_a, n_only_b, 0}.  Each section has the full per-atom capacity kcap
// (a section is never longer than the longer of the two rows), so the overflow rule stays "max row length > kcap".
//
// The build (k_build_pairs) is the packed-FP32 build of k_build_list_v3 with both atoms tested against every candidate
// record it loads (half the record loads per atom, same FP32 work), acceptance as four 32-candidate masks; the three
// sections are mask expressions (a & b, a & ~b, b & ~a) appended by walking set bits.  Semantics per atom are those of
// LJVPBuildListManager::build_neighbour_list (src/potentials/lennard_jones.rs:345-415) with rcut := rcut + skin.
#pragma once

namespace pisb {

struct PairListArgs {
    int npp;   // padded number of pair threads (multiple of 32)
    int kcap;  // capacity of EACH section (whole K-tiles of 4)
    int *both, *only_a, *only_b;
    int4 *counts;
};

__host__ __device__ __forceinline__ size_t plist_at(int k, int p, int npp) { return ((size_t)(k >> 2) * (size_t)npp + (size_t)p) * 4 + (size_t)(k & 3); }

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
template <bool IMAGE, int NPAIRS = 2>
__device__ __forceinline__ int build_pair_body(const Build2Args &a, const PairListArgs &pl, const float *__restrict__ xp, int p, bool act_a,
                                               bool act_b) {
    const int ia = 2 * p, ib = 2 * p + 1;
    const double4 xa = a.xt[ia];
    const double4 xb = act_b ? a.xt[ib] : xa;
    const float4 fa = a.xf[ia];
    const float4 fb = act_b ? a.xf[ib] : fa;
    int ca[3], cb[3];
    cell_coords<true>(a.box, a.g, xa.x, xa.y, xa.z, ca);
    cell_coords<true>(a.box, a.g, xb.x, xb.y, xb.z, cb);
    const float lo = a.pairf0.lo_list, hi = a.pairf0.hi_list;
    const float Lx = a.boxf.L[0], Ly = a.boxf.L[1], Lz = a.boxf.L[2];
    const float iLx = a.boxf.invL[0], iLy = a.boxf.invL[1], iLz = a.boxf.invL[2];
    const f32x2_t xa2 = f2_pack(fa.x, fa.x), ya2 = f2_pack(fa.y, fa.y), za2 = f2_pack(fa.z, fa.z);
    const f32x2_t xb2 = f2_pack(fb.x, fb.x), yb2 = f2_pack(fb.y, fb.y), zb2 = f2_pack(fb.z, fb.z);
    const f32x2_t magic = f2_pack(12582912.0f, 12582912.0f);
    auto image = [&](f32x2_t d, float L, float iL) {
        const f32x2_t t = f2_mul(d, f2_pack(iL, iL));
        const f32x2_t r = f2_sub(f2_add(t, magic), magic);
        return f2_fma(f2_neg(r), f2_pack(L, L), d);
    };
    // squared FP32 distances of one atom to both candidates of a record: the same IEEE operations as r2_f32
    auto dist2 = [&](const PairRec &r, f32x2_t x2, f32x2_t y2, f32x2_t z2, float &ra, float &rb) {
        f32x2_t dx = f2_sub(r.x, x2), dy = f2_sub(r.y, y2), dz = f2_sub(r.z, z2);
        if (IMAGE) {
            dx = image(dx, Lx, iLx);
            dy = image(dy, Ly, iLy);
            dz = image(dz, Lz, iLz);
        }
        f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), ra, rb);
    };
    int cnt_both = 0, cnt_a = 0, cnt_b = 0;
    int *wp_both = pl.both + (size_t)p * 4, *wp_a = pl.only_a + (size_t)p * 4, *wp_b = pl.only_b + (size_t)p * 4;
    const ptrdiff_t tile_step = (ptrdiff_t)pl.npp * 4 - 3;
    const int kcap = pl.kcap;
    auto emit = [&](unsigned m, int j0, int *&wp, int &cnt) {
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1u;
            if (cnt < kcap) *wp = j0 + bit;
            ++cnt;
            wp += (cnt & 3) ? (ptrdiff_t)1 : tile_step;
        }
    };
    auto exact_out = [&](const double4 &xi, int jj) {
        return build3_exact_out(a.box.h[0], a.box.h[4], a.box.h[8], a.box.hinv[0], a.box.hinv[4], a.box.hinv[8], a.xt, xi.x, xi.y, xi.z, jj,
                                a.pair0.t_list);
    };
    // candidates of the slot range [jb, je) against atom a (do_a) and / or atom b (do_b)
    auto scan = [&](int jb, int je, bool do_a, bool do_b) {
        const int q_end = (je + 1) >> 1;
        for (int q0 = jb >> 1; q0 < q_end; q0 += 16) {
            unsigned ha = 0u, la = 0u, hb = 0u, lb = 0u;  // bit c <-> candidate slot 2*q0 + c
#pragma unroll 1
            for (int s = 0; s < 16 && q0 + s < q_end; s += NPAIRS) {
                PairRec r[NPAIRS];
#pragma unroll
                for (int u = 0; u < NPAIRS; ++u) r[u] = ldg_pair(xp, q0 + s + u);  // array padded: reads past q_end stay in bounds
                unsigned bha = 0u, bla = 0u, bhb = 0u, blb = 0u;
#pragma unroll
                for (int u = 0; u < NPAIRS; ++u) {
                    float r0, r1;
                    dist2(r[u], xa2, ya2, za2, r0, r1);
                    if (r0 <= hi) bha |= 1u << (2 * u);
                    if (r1 <= hi) bha |= 2u << (2 * u);
                    if (r0 < lo) bla |= 1u << (2 * u);
                    if (r1 < lo) bla |= 2u << (2 * u);
                    dist2(r[u], xb2, yb2, zb2, r0, r1);
                    if (r0 <= hi) bhb |= 1u << (2 * u);
                    if (r1 <= hi) bhb |= 2u << (2 * u);
                    if (r0 < lo) blb |= 1u << (2 * u);
                    if (r1 < lo) blb |= 2u << (2 * u);
                }
                ha |= bha << (2 * s);
                la |= bla << (2 * s);
                hb |= bhb << (2 * s);
                lb |= blb << (2 * s);
            }
            // range edges (and whatever lies in the padding / beyond q_end), the atoms themselves
            const int first = jb - 2 * q0, last = je - 2 * q0;  // valid bits: [first, last)
            unsigned valid = last >= 32 ? 0xffffffffu : ((1u << last) - 1u);
            if (first > 0) valid &= ~((1u << first) - 1u);
            const unsigned self_a = (unsigned)(ia - 2 * q0) < 32u ? 1u << (ia - 2 * q0) : 0u;
            const unsigned self_b = (unsigned)(ib - 2 * q0) < 32u ? 1u << (ib - 2 * q0) : 0u;
            ha = do_a ? ha & valid & ~self_a : 0u;
            hb = do_b ? hb & valid & ~self_b : 0u;
            // FP32 guard band: the exact FP64 reference predicate decides
            unsigned band = ha & ~la;
            while (band) {
                const int bit = __ffs(band) - 1;
                band &= band - 1u;
                if (exact_out(xa, 2 * q0 + bit)) ha &= ~(1u << bit);
            }
            band = hb & ~lb;
            while (band) {
                const int bit = __ffs(band) - 1;
                band &= band - 1u;
                if (exact_out(xb, 2 * q0 + bit)) hb &= ~(1u << bit);
            }
            emit(ha & hb, 2 * q0, wp_both, cnt_both);
            emit(ha & ~hb, 2 * q0, wp_a, cnt_a);
            emit(hb & ~ha, 2 * q0, wp_b, cnt_b);
        }
    };
    const int nx = a.g.n[0];
    // the stencil rows of cell row (c[1], c[2]) over the x-cells [cx_lo + lo, cx_hi + hi]
    auto rows = [&](const int *c, int cx_lo, int cx_hi, bool do_a, bool do_b) {
        for (int dz = a.g.lo[2]; dz <= a.g.hi[2]; ++dz) {
            int cz = c[2] + dz;
            if (cz < 0 || cz >= a.g.n[2]) {
                // interior atoms: a wrapped cell lies beyond the cutoff (margin > rc + skin); brick-local dims never wrap
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
                const int xlo = cx_lo + a.g.lo[0], xhi = cx_hi + a.g.hi[0];
                if (IMAGE && xlo < 0 && !a.g.local[0]) scan(__ldg(&a.cell_start[rb + xlo + nx]), __ldg(&a.cell_start[rb + nx]), do_a, do_b);
                scan(__ldg(&a.cell_start[rb + max(xlo, 0)]), __ldg(&a.cell_start[rb + min(xhi, nx - 1) + 1]), do_a, do_b);
                if (IMAGE && xhi >= nx && !a.g.local[0]) scan(__ldg(&a.cell_start[rb]), __ldg(&a.cell_start[rb + xhi - nx + 1]), do_a, do_b);
            }
        }
    };
    // One joint sweep when the two atoms sit in the same cell row, at most two cells apart, and the widened x-range does not
    // lap the periodic row; otherwise (row ends, vacuum gaps, tiny grids, a ghost partner) each atom gets its own sweep and
    // BOTH stays empty -- same sets either way.
    const bool joint = act_a && act_b && ca[1] == cb[1] && ca[2] == cb[2] && cb[0] >= ca[0] && cb[0] - ca[0] <= 2 &&
                       (a.g.hi[0] - a.g.lo[0] + 1 + cb[0] - ca[0]) <= nx;
    if (joint) {
        rows(ca, ca[0], cb[0], true, true);
    } else {
        if (act_a) rows(ca, ca[0], ca[0], true, false);
        if (act_b) rows(cb, cb[0], cb[0], false, true);
    }
    pl.counts[p] = make_int4(min(cnt_both, kcap), min(cnt_a, kcap), min(cnt_b, kcap), 0);
    a.nnbr[ia] = act_a ? cnt_both + cnt_a : 0;
    if (ib < a.n) a.nnbr[ib] = act_b ? cnt_both + cnt_b : 0;
    return max(cnt_both + cnt_a, cnt_both + cnt_b);
}

// two atoms' worth of masks and cursors: 80 registers (6 blocks of 128 pair threads = 1536 atoms per SM in flight)
__global__ void __launch_bounds__(TPB_FORCE, 6) k_build_pairs(Build2Args a, PairListArgs pl, const float *__restrict__ xp) {
    if (a.flags[FLAG_REBUILD] == 0) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = 2 * p, ib = 2 * p + 1;
    bool act_a = false, act_b = false, interior = true;
    if (ia < a.n) {
        const float4 fa = a.xf[ia];
        act_a = !xf_is_ghost(fa);
        if (act_a) interior = is_interior(a.boxf, fa);
        if (ib < a.n) {
            const float4 fb = a.xf[ib];
            act_b = !xf_is_ghost(fb);
            if (act_b) interior = interior && is_interior(a.boxf, fb);
        }
    }
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    int m = 0;
    if (act_a || act_b) {
        // a ghost partner (multi-GPU) is simply never swept for: the body reads its record (a valid slot) but lists nothing for it
        m = warp_interior ? build_pair_body<false>(a, pl, xp, p, act_a, act_b) : build_pair_body<true>(a, pl, xp, p, act_a, act_b);
    } else if (ia < a.n) {
        pl.counts[p] = make_int4(0, 0, 0, 0);
        a.nnbr[ia] = 0;
        if (ib < a.n) a.nnbr[ib] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&a.flags[FLAG_MAXNBR], m);
}

// ------------------------------------------------------------------------------------------------
// force
// ------------------------------------------------------------------------------------------------
// sums of a pair thread: forces of both atoms, ONE set of energy sums (PE and virial are totals anyway)
struct PairAcc {
    double fax = 0.0, fay = 0.0, faz = 0.0, fbx = 0.0, fby = 0.0, fbz = 0.0;
    double s12 = 0.0, s6 = 0.0;
    int cnt = 0;
    __device__ __forceinline__ void pair(const PairDev &p, double dx, double dy, double dz, double r2, double &fx, double &fy, double &fz) {
        const double y = rcp_newton(r2);
        const double s2 = p.sig2 * y;
        const double q6 = (s2 * s2) * s2;
        const double q12 = q6 * q6;
        const double w = fma(2.0, q12, -q6) * y;
        fx = fma(w, dx, fx);
        fy = fma(w, dy, fy);
        fz = fma(w, dz, fz);
        s12 += q12;
        s6 += q6;
        ++cnt;
    }
};

// one section of the pair list: MASK bit 0 -> evaluate against atom a, bit 1 -> against atom b
template <bool IMAGE, int MASK>
__device__ __forceinline__ void pair_section(const double4 *__restrict__ xt, const BoxDev &box, const PairDev &p0, const int4 *__restrict__ tiles,
                                             int npp, int cnt, const double4 &xa, const double4 &xb, PairAcc &acc, bool &ambiguous) {
    int4 cur = cnt > 0 ? ldg_stream_i4(tiles) : make_int4(0, 0, 0, 0);
    for (int k = 0; k < cnt; k += 4) {
        int4 nxt = cur;
        if (k + 4 < cnt) nxt = ldg_stream_i4(tiles + (size_t)((k >> 2) + 1) * npp);
        int j[4] = {cur.x, cur.y, cur.z, cur.w};
        bool in[4];
        double4 xj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            in[u] = k + u < cnt;
            if (!in[u]) j[u] = 0;  // tail of the last tile: any valid slot
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = ldg_d4(&xt[j[u]]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MASK & 1) {
                double dx, dy, dz;
                const double r2 = lean_disp<IMAGE>(box, xa, xj[u], dx, dy, dz);
                if (in[u]) {
                    if (le_bits(r2, p0.t_lo)) acc.pair(p0, dx, dy, dz, r2, acc.fax, acc.fay, acc.faz);
                    else ambiguous |= le_bits(r2, p0.t_hi);
                }
            }
            if (MASK & 2) {
                double dx, dy, dz;
                const double r2 = lean_disp<IMAGE>(box, xb, xj[u], dx, dy, dz);
                if (in[u]) {
                    if (le_bits(r2, p0.t_lo)) acc.pair(p0, dx, dy, dz, r2, acc.fbx, acc.fby, acc.fbz);
                    else ambiguous |= le_bits(r2, p0.t_hi);
                }
            }
        }
        cur = nxt;
    }
}

// A pair thread that met a guard-band pair (one in ~10^11) is redone here with the reference-order predicate available;
// out of line and after the loops (see force_atom_exact).  out9 = fa(3), fb(3), s12, s6, cnt.
template <bool IMAGE>
__device__ __noinline__ void pair_thread_exact(const double4 *__restrict__ xt, const int *__restrict__ both, const int *__restrict__ only_a,
                                               const int *__restrict__ only_b, int npp, int p, int4 cnt, BoxDev box, PairDev p0, bool has_b,
                                               double *out9) {
    const double4 xa = xt[2 * p];
    const double4 xb = has_b ? xt[2 * p + 1] : xa;
    PairAcc acc;
    for (int sec = 0; sec < 3; ++sec) {
        const int *base = sec == 0 ? both : (sec == 1 ? only_a : only_b);
        const int n = sec == 0 ? cnt.x : (sec == 1 ? cnt.y : cnt.z);
        for (int k = 0; k < n; ++k) {
            const double4 xj = ldg_d4(&xt[base[plist_at(k, p, npp)]]);
            double dx, dy, dz;
            if (sec != 2) {
                const double r2 = lean_disp<IMAGE>(box, xa, xj, dx, dy, dz);
                if (lean_in_range(box, p0, xa, xj, r2)) acc.pair(p0, dx, dy, dz, r2, acc.fax, acc.fay, acc.faz);
            }
            if (sec != 1) {
                const double r2 = lean_disp<IMAGE>(box, xb, xj, dx, dy, dz);
                if (lean_in_range(box, p0, xb, xj, r2)) acc.pair(p0, dx, dy, dz, r2, acc.fbx, acc.fby, acc.fbz);
            }
        }
    }
    out9[0] = acc.fax, out9[1] = acc.fay, out9[2] = acc.faz, out9[3] = acc.fbx, out9[4] = acc.fby, out9[5] = acc.fbz;
    out9[6] = acc.s12, out9[7] = acc.s6, out9[8] = (double)acc.cnt;
}

template <bool IMAGE>
__device__ __forceinline__ void pair_thread_body(const Force2Args &a, const PairListArgs &pl, int p, bool has_b, PairAcc &acc) {
    const double4 xa = a.xt[2 * p];
    const double4 xb = has_b ? a.xt[2 * p + 1] : xa;
    const int4 cnt = pl.counts[p];
    bool ambiguous = false;
    pair_section<IMAGE, 3>(a.xt, a.box, a.pair0, reinterpret_cast<const int4 *>(pl.both) + p, pl.npp, cnt.x, xa, xb, acc, ambiguous);
    pair_section<IMAGE, 1>(a.xt, a.box, a.pair0, reinterpret_cast<const int4 *>(pl.only_a) + p, pl.npp, cnt.y, xa, xb, acc, ambiguous);
    pair_section<IMAGE, 2>(a.xt, a.box, a.pair0, reinterpret_cast<const int4 *>(pl.only_b) + p, pl.npp, cnt.z, xa, xb, acc, ambiguous);
    if (ambiguous) {
        double o[9];
        pair_thread_exact<IMAGE>(a.xt, pl.both, pl.only_a, pl.only_b, pl.npp, p, cnt, a.box, a.pair0, has_b, o);
        acc.fax = o[0], acc.fay = o[1], acc.faz = o[2], acc.fbx = o[3], acc.fby = o[4], acc.fbz = o[5];
        acc.s12 = o[6], acc.s6 = o[7], acc.cnt = (int)o[8];
    }
}

// kick of step k (+ KE, tr(X F^T)) and, when DRIFT, drift + wrap + skin trigger of step k+1 for ONE atom: the epilogue of
// k_force_vv, operation for operation (potential.rs:16-22, :28-30).  red[2..5] accumulate.
template <bool DRIFT>
__device__ __forceinline__ void vv_epilogue_atom(const ForceVVArgs &b, int i, double fx, double fy, double fz, double *red) {
    const Force2Args &a = b.f;
    double4 x = a.xt[i];
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
    red[2] += __dmul_rn(__dmul_rn(0.5, m), norm2(vx, vy, vz));
    red[3] += __dmul_rn(x.x, fx);
    red[4] += __dmul_rn(x.y, fy);
    red[5] += __dmul_rn(x.z, fz);
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

// k_pforce<FUSED, DRIFT, BRICK>: the force pass of a pair-list system.
//   FUSED = false : LJVOffsetManager::compute_potential (lennard_jones.rs:186-244): F (+ accumulate source), PE, pair virial
//   FUSED = true  : + the integrator epilogue of k_force_vv (one launch per NVE step); DRIFT as there
//   BRICK         : multi-GPU -- ghost slots get no forces and the launch may be speculative (skip_flag)
// 5 resident blocks of 128 pair threads per SM (<= 96 registers): 1280 atoms in flight per SM.
template <bool FUSED, bool DRIFT, bool BRICK>
__global__ void __launch_bounds__(TPB_FORCE, 5) k_pforce(ForceVVArgs b, PairListArgs pl) {
    const Force2Args &a = b.f;
    if (BRICK && a.skip_flag && *a.skip_flag != 0) return;  // speculative launch, a rebuild comes first
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = 2 * p, ib = 2 * p + 1;
    double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // pe, pair virial, ke, x*fx, y*fy, z*fz
    if (FUSED && DRIFT && p == 0) b.flags[b.unwrapped_out] = 0;
    bool act_a = false, act_b = false;
    bool interior = *a.unwrapped == 0;
    if (ia < a.n) {
        const float4 fa = a.xf[ia];
        act_a = !(BRICK && xf_is_ghost(fa));
        if (act_a) interior = interior && is_interior(a.boxf, fa);
        if (ib < a.n) {
            const float4 fb = a.xf[ib];
            act_b = !(BRICK && xf_is_ghost(fb));
            if (act_b) interior = interior && is_interior(a.boxf, fb);
        }
    }
    const bool warp_interior = __all_sync(0xffffffffu, interior);
    if (act_a || act_b) {
        PairAcc acc;
        if (warp_interior) pair_thread_body<false>(a, pl, p, ib < a.n, acc);
        else pair_thread_body<true>(a, pl, p, ib < a.n, acc);
        const PairDev &p0 = a.pair0;
        red[0] = fma(p0.c4, acc.s12 - acc.s6, -(double)acc.cnt * p0.ucut);
        red[1] = p0.c24 * fma(2.0, acc.s12, -acc.s6);
        double fx = -p0.c24 * acc.fax, fy = -p0.c24 * acc.fay, fz = -p0.c24 * acc.faz;
        if (act_a) {
            if (!FUSED && a.ax) {
                fx += a.ax[ia];
                fy += a.ay[ia];
                fz += a.az[ia];
            }
            a.fx[ia] = fx;
            a.fy[ia] = fy;
            a.fz[ia] = fz;
            if (FUSED) vv_epilogue_atom<DRIFT>(b, ia, fx, fy, fz, red);
        }
        if (act_b) {
            fx = -p0.c24 * acc.fbx, fy = -p0.c24 * acc.fby, fz = -p0.c24 * acc.fbz;
            if (!FUSED && a.ax) {
                fx += a.ax[ib];
                fy += a.ay[ib];
                fz += a.az[ib];
            }
            a.fx[ib] = fx;
            a.fy[ib] = fy;
            a.fz[ib] = fz;
            if (FUSED) vv_epilogue_atom<DRIFT>(b, ib, fx, fy, fz, red);
        }
    }
    pisb_thermo *th = a.thermo;
    if (FUSED) {
        double t3[3];
        block_reduce_finalize<6, TPB_FORCE>(red, a.partials, a.ticket, [&](int q, double s) {
            if (q == 0) th->pe = s / 2.0;
            else if (q == 1) th->virial_pair = s / 2.0;
            else if (q == 2) th->ke = s;
            else t3[q - 3] = s;
            if (q == 5) th->virial_ref = (t3[0] + t3[1]) + t3[2];
        });
    } else {
        double r2[2] = {red[0], red[1]};
        block_reduce_finalize<2, TPB_FORCE>(r2, a.partials, a.ticket, [&](int q, double s) {
            if (q == 0) th->pe = s / 2.0;
            else th->virial_pair = s / 2.0;
        });
    }
}

// ------------------------------------------------------------------------------------------------
// pisb_list_stats: what the CURRENT list holds, counted on the device from whichever list form is in use --
// listed pairs (sum of the per-atom row lengths), pairs inside the cutoff right now (the reference predicate), index
// words stored (pair lists store the common neighbours of two atoms once).  bench.py's roofline needs K and K_in.
// ------------------------------------------------------------------------------------------------
struct ListStatsArgs {
    int n, npad, pairs;
    const double4 *xt;
    const int *nbr, *nnbr;
    PairListArgs pl;
    BoxDev box;
    PairDev pair0;
    const PairDev *table;
    int n_types;
    unsigned long long *out;  // [0] listed, [1] in range, [2] index words
};

template <bool ORTHO>
__global__ void __launch_bounds__(TPB) k_list_stats(ListStatsArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long listed = 0, inr = 0, words = 0;
    if (i < a.n && !is_ghost(a.xt[i].w)) {
        const double4 xi = a.xt[i];
        const int ti = type_of(xi.w);
        const int nn = a.nnbr[i];
        const int p = i >> 1;
        const int4 c = a.pairs ? a.pl.counts[p] : make_int4(0, 0, 0, 0);
        listed = (unsigned long long)nn;
        // an atom's share of the stored words: its own section + BOTH once per pair thread (charged to the even slot, or to
        // the odd one when the even slot is a ghost and lists nothing)
        if (a.pairs) words = (unsigned long long)((i & 1) ? c.z : c.y) + (((i & 1) == 0 || is_ghost(a.xt[i - 1].w)) ? (unsigned long long)c.x : 0ull);
        else words = (unsigned long long)nn;
        for (int k = 0; k < nn; ++k) {
            int j;
            if (!a.pairs) j = a.nbr[nbr_at(k, i, a.npad)];
            else if (k < c.x) j = a.pl.both[plist_at(k, p, a.pl.npp)];
            else j = ((i & 1) ? a.pl.only_b : a.pl.only_a)[plist_at(k - c.x, p, a.pl.npp)];
            const double4 xj = ldg_d4(&a.xt[j]);
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
        words += __shfl_down_sync(0xffffffffu, words, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.out[0], listed);
        atomicAdd(&a.out[1], inr);
        atomicAdd(&a.out[2], words);
    }
}

}  // namespace pisb
