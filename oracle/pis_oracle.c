/*
 * pis_oracle.c -- CPU ORACLE for the PIS per-timestep MD hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (pis_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference (vivekadishankara/pis, Rust) cannot be
 * compiled in this image (no rustc/cargo, no vendored crates) and its own tests
 * (src/tests/command_tests.rs) pin no force, energy, cell, list or integrator result.
 * This file restates the reference arithmetic operation-for-operation (operation order,
 * no FMA contraction: build with -ffp-contract=off) and is pinned instead against
 * analytic known answers and an independent numpy O(N^2) restatement (tests/test_oracle.py).
 *
 * Third-party arithmetic restated from its published algorithm (source is not under
 * /root/reference): nalgebra 0.34.1 (Cargo.lock:494-495)
 *   - Matrix3::try_inverse  : adjugate / determinant (linalg/inverse.rs, 3x3 arm)
 *   - Matrix3 * Vector3     : column-axpy gemv, y_i = ((a_i0*x0) + a_i1*x1) + a_i2*x2
 *   - Vector3::norm_squared : (x*x + y*y) + z*z   (base/blas.rs dot, 3-vector arm)
 *   - Matrix3::determinant  : m11*(m22*m33-m32*m23) - m12*(m21*m33-m31*m23) + m13*(m21*m32-m31*m22)
 *   - Matrix3xX * Matrix3xX^T (virial): per output column, sequential gemv over atoms.
 * f64::powi(n) for n = 2, 3 is repeated multiplication ((x*x)*x).
 *
 * All `ref:` citations are paths relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* Box: h and h_inv are column-major 3x3 like nalgebra (element (r,c) at [c*3+r]).
 * ref: src/simulation_box.rs:5-9 */
typedef struct {
    double h[9];
    double hinv[9];
    int pbc[3];
} orc_box;

/* Pair table: dense n_types x n_types, entry (i,j) 1-based at [(i-1)*n_types + (j-1)],
 * stored AS GIVEN (not symmetrised); lookups read (min,max) like the reference's sorted
 * HashMap key.  present[] == 0 means no entry.  ref: src/potentials/potential.rs:143-144,181-192 */
typedef struct {
    int n_types;
    const double *eps;
    const double *sigma;
    const double *rcut;
    const unsigned char *present;
    int shift; /* ref: lennard_jones.rs:18,44 -- always true at every construction site */
} orc_table;

#define H(b, r, c) ((b)->h[(c) * 3 + (r)])
#define HI(b, r, c) ((b)->hinv[(c) * 3 + (r)])

/* nalgebra Matrix3::determinant */
static double det3(const double *m) {
#define M(r, c) m[((c) - 1) * 3 + ((r) - 1)]
    double minor_m12_m23 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    double minor_m11_m23 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    double minor_m11_m22 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    return M(1, 1) * minor_m12_m23 - M(1, 2) * minor_m11_m23 + M(1, 3) * minor_m11_m22;
}

/* SimulationBox::new: h_inv = h.try_inverse().  ref: src/simulation_box.rs:12-15.
 * Returns 0 on success, 1 if singular ("Box matrix should be invertible"). */
ORC_API int orc_box_new(const double *h9, const int *pbc3, orc_box *out) {
    const double *m = h9;
    double minor_m12_m23 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    double minor_m11_m23 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    double minor_m11_m22 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    double determinant =
        M(1, 1) * minor_m12_m23 - M(1, 2) * minor_m11_m23 + M(1, 3) * minor_m11_m22;
    if (determinant == 0.0) return 1;
    memcpy(out->h, h9, sizeof(double) * 9);
    double *o = out->hinv;
#define O(r, c) o[((c) - 1) * 3 + ((r) - 1)]
    O(1, 1) = minor_m12_m23 / determinant;
    O(1, 2) = (M(1, 3) * M(3, 2) - M(3, 3) * M(1, 2)) / determinant;
    O(1, 3) = (M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3)) / determinant;
    O(2, 1) = -minor_m11_m23 / determinant;
    O(2, 2) = (M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3)) / determinant;
    O(2, 3) = (M(1, 3) * M(2, 1) - M(2, 3) * M(1, 1)) / determinant;
    O(3, 1) = minor_m11_m22 / determinant;
    O(3, 2) = (M(1, 2) * M(3, 1) - M(3, 2) * M(1, 1)) / determinant;
    O(3, 3) = (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) / determinant;
#undef O
#undef M
    for (int d = 0; d < 3; ++d) out->pbc[d] = pbc3[d];
    return 0;
}

/* SimulationBox::from_lammps_data.  ref: src/simulation_box.rs:44-65 (pbc always true) */
ORC_API int orc_box_from_lammps(double xlo, double xhi, double ylo, double yhi, double zlo,
                                double zhi, double xy, double xz, double yz, orc_box *out) {
    double h[9] = {xhi - xlo, 0.0, 0.0, xy, yhi - ylo, 0.0, xz, yz, zhi - zlo};
    int pbc[3] = {1, 1, 1};
    return orc_box_new(h, pbc, out);
}

/* SimulationBox::volume.  ref: src/simulation_box.rs:67-69 */
ORC_API double orc_box_volume(const orc_box *b) { return fabs(det3(b->h)); }

/* nalgebra Matrix3 * Vector3 (column-axpy order) */
static inline void matvec3(const double *m, const double *x, double *y) {
    double y0 = m[0] * x[0], y1 = m[1] * x[0], y2 = m[2] * x[0];
    y0 = m[3] * x[1] + y0;
    y1 = m[4] * x[1] + y1;
    y2 = m[5] * x[1] + y2;
    y0 = m[6] * x[2] + y0;
    y1 = m[7] * x[2] + y1;
    y2 = m[8] * x[2] + y2;
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
}

/* apply_boundary_conditions_dis (minimum image).  ref: src/simulation_box.rs:17-27 */
ORC_API void orc_min_image(const orc_box *b, double *rij) {
    double s[3];
    matvec3(b->hinv, rij, s);
    for (int d = 0; d < 3; ++d)
        if (b->pbc[d]) s[d] -= round(s[d]); /* Rust f64::round: half away from zero */
    matvec3(b->h, s, rij);
}

/* apply_boundary_conditions_pos (wrap).  ref: src/simulation_box.rs:29-42 */
ORC_API void orc_wrap_pos(const orc_box *b, double *r) {
    double s[3];
    matvec3(b->hinv, r, s);
    for (int d = 0; d < 3; ++d)
        if (b->pbc[d]) s[d] -= floor(s[d]);
    matvec3(b->h, s, r);
}

static inline double norm_squared3(const double *v) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }

/* Rust `f64 as usize`: saturating, NaN -> 0 */
static inline uint64_t f64_as_usize(double x) {
    if (!(x > 0.0)) return 0; /* negative, -0, NaN */
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}

/* PairPotentialManager::max_rcut.  ref: src/potentials/potential.rs:170-179 */
ORC_API double orc_max_rcut(const orc_table *t) {
    double max_rcut = 0.0;
    for (int k = 0; k < t->n_types * t->n_types; ++k)
        if (t->present[k] && max_rcut < t->rcut[k]) max_rcut = t->rcut[k];
    return max_rcut;
}

/* Atoms::divide_into_cells.  ref: src/atoms/neighbour_list.rs:45-57 */
ORC_API void orc_divide_into_cells(const orc_box *b, double rcut, uint64_t *ncell) {
    for (int i = 0; i < 3; ++i) {
        const double *col = &b->h[i * 3];
        double len_i = sqrt(norm_squared3(col));
        ncell[i] = f64_as_usize(floor(len_i / rcut));
        if (ncell[i] == 0) ncell[i] = 1;
    }
}

/* Atoms::rcut_cells + cell_index, as CSR (push order == ascending atom index).
 * ref: src/atoms/neighbour_list.rs:9-42,61-63.  cell_of[n], cell_start[ncells+1], cell_atoms[n] */
ORC_API void orc_rcut_cells(const orc_box *b, const double *pos, int64_t n, uint64_t nx,
                            uint64_t ny, uint64_t nz, int64_t *cell_of, int64_t *cell_start,
                            int64_t *cell_atoms) {
    int64_t ncell_total = (int64_t)(nx * ny * nz);
    memset(cell_start, 0, sizeof(int64_t) * (size_t)(ncell_total + 1));
    for (int64_t i = 0; i < n; ++i) {
        double s[3];
        matvec3(b->hinv, &pos[3 * i], s);
        uint64_t cx = f64_as_usize(floor(s[0] * (double)nx));
        uint64_t cy = f64_as_usize(floor(s[1] * (double)ny));
        uint64_t cz = f64_as_usize(floor(s[2] * (double)nz));
        if (b->pbc[0]) cx = cx % nx;
        if (b->pbc[1]) cy = cy % ny;
        if (b->pbc[2]) cz = cz % nz;
        /* non-periodic out-of-range would panic (OOB) in the reference; clamp is never hit: pbc is
           always true from the reader (commands.rs:352,362 -> simulation_box.rs:64) */
        int64_t c = (int64_t)((cz * ny + cy) * nx + cx);
        cell_of[i] = c;
        cell_start[c + 1]++;
    }
    for (int64_t c = 0; c < ncell_total; ++c) cell_start[c + 1] += cell_start[c];
    int64_t *fill = (int64_t *)calloc((size_t)ncell_total, sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = cell_of[i];
        cell_atoms[cell_start[c] + fill[c]++] = i;
    }
    free(fill);
}

/* FORWARD_NEIGHBOUR_OFFSETS.  ref: src/atoms/neighbour_list.rs:94-165 */
static const int FWD[14][3] = {{0, 0, 0},  {1, 0, 0},  {-1, 1, 0}, {0, 1, 0},  {1, 1, 0},
                               {-1, -1, 1}, {0, -1, 1}, {1, -1, 1}, {-1, 0, 1}, {0, 0, 1},
                               {1, 0, 1},  {-1, 1, 1}, {0, 1, 1},  {1, 1, 1}};
ORC_API const int *orc_forward_offsets(void) { return &FWD[0][0]; }

/* LennardJones::compute_potential.  ref: src/potentials/lennard_jones.rs:33-55.
 * f is the force on j (i receives -f). */
ORC_API void orc_lj_pair(double epsilon, double sigma, double rcut, int shift, const double *rij,
                         double *u_out, double *f) {
    double rij2 = norm_squared3(rij);
    double inv_rij2 = 1.0 / rij2;
    double s2 = (sigma * sigma) * inv_rij2;
    double vanderwaals_attraction = (s2 * s2) * s2; /* powi(3) */
    double lj_repulsion = vanderwaals_attraction * vanderwaals_attraction;
    double potential_energy = 4.0 * epsilon * (lj_repulsion - vanderwaals_attraction);
    double fs = 24.0 * epsilon * (2.0 * lj_repulsion - vanderwaals_attraction) * inv_rij2;
    f[0] = fs * rij[0];
    f[1] = fs * rij[1];
    f[2] = fs * rij[2];
    if (shift) {
        double q = sigma / rcut;
        double cutoff_inv2 = q * q;
        double cutoff_attraction = (cutoff_inv2 * cutoff_inv2) * cutoff_inv2;
        double cutoff_repulsion = cutoff_attraction * cutoff_attraction;
        double u_cutoff = 4.0 * epsilon * (cutoff_repulsion - cutoff_attraction);
        potential_energy -= u_cutoff;
    }
    *u_out = potential_energy;
}

/* get_potential_ij: sorted 1-based type pair.  ref: src/potentials/potential.rs:181-192 */
static inline int table_lookup(const orc_table *t, int type_i, int type_j) {
    int a = type_i < type_j ? type_i : type_j;
    int b = type_i < type_j ? type_j : type_i;
    if (a < 1 || b > t->n_types) return -1;
    int k = (a - 1) * t->n_types + (b - 1);
    return t->present[k] ? k : -1;
}

static inline int64_t rem_euclid(int64_t a, int64_t n) {
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

static int g_quiet_missing = 0;
ORC_API void orc_set_quiet(int q) { g_quiet_missing = q; }

/* LJVOffsetManager::compute_potential -- THE LIVE FORCE DRIVER, faithful serial restatement.
 * ref: src/potentials/lennard_jones.rs:186-244.  Accumulates into forces (caller zeroes);
 * returns total shifted PE.  rcut_cells_override <= 0 -> use max_rcut (the reference behaviour). */
ORC_API double orc_compute_potential(const orc_box *b, const orc_table *t, int64_t n,
                                     const double *pos, const int32_t *types, double *forces) {
    double max_rcut = orc_max_rcut(t);
    uint64_t nc[3];
    orc_divide_into_cells(b, max_rcut, nc);
    int64_t nx = (int64_t)nc[0], ny = (int64_t)nc[1], nz = (int64_t)nc[2];
    int64_t ncell_total = nx * ny * nz;
    int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cell_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncell_total + 1));
    int64_t *cell_atoms = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_rcut_cells(b, pos, n, nc[0], nc[1], nc[2], cell_of, cell_start, cell_atoms);

    double potential_energy = 0.0;
    for (int64_t cx_i = 0; cx_i < nx; ++cx_i)
        for (int64_t cy_i = 0; cy_i < ny; ++cy_i)
            for (int64_t cz_i = 0; cz_i < nz; ++cz_i) {
                int64_t current_cell = (cz_i * ny + cy_i) * nx + cx_i;
                for (int o = 0; o < 14; ++o) {
                    int64_t cx_j = rem_euclid(cx_i + FWD[o][0], nx);
                    int64_t cy_j = rem_euclid(cy_i + FWD[o][1], ny);
                    int64_t cz_j = rem_euclid(cz_i + FWD[o][2], nz);
                    int64_t a_neighbour_cell = (cz_j * ny + cy_j) * nx + cx_j;
                    for (int64_t ii = cell_start[current_cell]; ii < cell_start[current_cell + 1]; ++ii) {
                        int64_t i = cell_atoms[ii];
                        for (int64_t jj = cell_start[a_neighbour_cell];
                             jj < cell_start[a_neighbour_cell + 1]; ++jj) {
                            int64_t j = cell_atoms[jj];
                            if (current_cell == a_neighbour_cell && i <= j) continue;
                            double rij[3] = {pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1],
                                             pos[3 * j + 2] - pos[3 * i + 2]};
                            orc_min_image(b, rij);
                            int k = table_lookup(t, types[i], types[j]);
                            if (k < 0) {
                                if (!g_quiet_missing)
                                    printf("During force calculation between %lld and %lld atoms, potential was missing\n",
                                           (long long)(i + 1), (long long)(j + 1));
                                continue;
                            }
                            if (sqrt(norm_squared3(rij)) > t->rcut[k]) continue;
                            double u, f[3];
                            orc_lj_pair(t->eps[k], t->sigma[k], t->rcut[k], t->shift, rij, &u, f);
                            potential_energy += u;
                            forces[3 * i] -= f[0];
                            forces[3 * i + 1] -= f[1];
                            forces[3 * i + 2] -= f[2];
                            forces[3 * j] += f[0];
                            forces[3 * j + 1] += f[1];
                            forces[3 * j + 2] += f[2];
                        }
                    }
                }
            }
    free(cell_of);
    free(cell_start);
    free(cell_atoms);
    return potential_energy;
}

/* All-core variant of the same loop: what LJVParallelManager (ref: lennard_jones.rs:248-338)
 * intends -- cells distributed over threads, per-thread force arrays reduced at the end.
 * Same per-pair arithmetic; only the summation order differs.  Used as the timed CPU baseline. */
ORC_API double orc_compute_potential_omp(const orc_box *b, const orc_table *t, int64_t n,
                                         const double *pos, const int32_t *types, double *forces,
                                         int n_threads) {
    double max_rcut = orc_max_rcut(t);
    uint64_t nc[3];
    orc_divide_into_cells(b, max_rcut, nc);
    int64_t nx = (int64_t)nc[0], ny = (int64_t)nc[1], nz = (int64_t)nc[2];
    int64_t ncell_total = nx * ny * nz;
    int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cell_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncell_total + 1));
    int64_t *cell_atoms = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_rcut_cells(b, pos, n, nc[0], nc[1], nc[2], cell_of, cell_start, cell_atoms);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    double *tf = (double *)calloc((size_t)n_threads * 3 * (size_t)(n > 0 ? n : 1), sizeof(double));
    double potential_energy = 0.0;
#pragma omp parallel num_threads(n_threads) reduction(+ : potential_energy)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *F = tf + (size_t)tid * 3 * (size_t)n;
#pragma omp for schedule(dynamic, 8)
        for (int64_t cc = 0; cc < ncell_total; ++cc) {
            int64_t cx_i = cc % nx, cy_i = (cc / nx) % ny, cz_i = cc / (nx * ny);
            int64_t current_cell = cc;
            for (int o = 0; o < 14; ++o) {
                int64_t cx_j = rem_euclid(cx_i + FWD[o][0], nx);
                int64_t cy_j = rem_euclid(cy_i + FWD[o][1], ny);
                int64_t cz_j = rem_euclid(cz_i + FWD[o][2], nz);
                int64_t a_neighbour_cell = (cz_j * ny + cy_j) * nx + cx_j;
                for (int64_t ii = cell_start[current_cell]; ii < cell_start[current_cell + 1]; ++ii) {
                    int64_t i = cell_atoms[ii];
                    for (int64_t jj = cell_start[a_neighbour_cell]; jj < cell_start[a_neighbour_cell + 1]; ++jj) {
                        int64_t j = cell_atoms[jj];
                        if (current_cell == a_neighbour_cell && i <= j) continue;
                        double rij[3] = {pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1],
                                         pos[3 * j + 2] - pos[3 * i + 2]};
                        orc_min_image(b, rij);
                        int k = table_lookup(t, types[i], types[j]);
                        if (k < 0) continue;
                        if (sqrt(norm_squared3(rij)) > t->rcut[k]) continue;
                        double u, f[3];
                        orc_lj_pair(t->eps[k], t->sigma[k], t->rcut[k], t->shift, rij, &u, f);
                        potential_energy += u;
                        F[3 * i] -= f[0];
                        F[3 * i + 1] -= f[1];
                        F[3 * i + 2] -= f[2];
                        F[3 * j] += f[0];
                        F[3 * j + 1] += f[1];
                        F[3 * j + 2] += f[2];
                    }
                }
            }
        }
#pragma omp for schedule(static)
        for (int64_t k = 0; k < 3 * n; ++k) {
            double s = 0.0;
            for (int th = 0; th < n_threads; ++th) s += tf[(size_t)th * 3 * (size_t)n + (size_t)k];
            forces[k] += s;
        }
    }
    free(tf);
    free(cell_of);
    free(cell_start);
    free(cell_atoms);
    return potential_energy;
}

/* LJManager::compute_potential -- O(N^2) all-pairs driver (ref: lennard_jones.rs:66-99).
 * An independent traversal of the same pair arithmetic; used to cross-check the cell driver. */
ORC_API double orc_compute_potential_n2(const orc_box *b, const orc_table *t, int64_t n,
                                        const double *pos, const int32_t *types, double *forces) {
    double potential_energy = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            double rij[3] = {pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1],
                             pos[3 * j + 2] - pos[3 * i + 2]};
            orc_min_image(b, rij);
            int k = table_lookup(t, types[i], types[j]);
            if (k < 0) continue;
            if (sqrt(norm_squared3(rij)) > t->rcut[k]) continue;
            double u, f[3];
            orc_lj_pair(t->eps[k], t->sigma[k], t->rcut[k], t->shift, rij, &u, f);
            potential_energy += u;
            forces[3 * i] -= f[0];
            forces[3 * i + 1] -= f[1];
            forces[3 * i + 2] -= f[2];
            forces[3 * j] += f[0];
            forces[3 * j + 1] += f[1];
            forces[3 * j + 2] += f[2];
        }
    return potential_energy;
}

/* LJVPBuildListManager::build_neighbour_list -- full list (i->j and j->i), 27-cell stencil,
 * per-dx-slab dedupe of wrapped cells, skip i==j, `norm() > rcut -> skip`.
 * ref: src/potentials/lennard_jones.rs:345-415.
 * `extra` is added to every pair rcut AND to the cell-grid rcut: extra = 0 is the reference;
 * extra = skin gives "reference predicate with rcut := rc + skin" (the Verlet-list oracle).
 * Output CSR: nbr_start[n+1]; nbr (capacity cap).  Returns total entries, or -1 if cap too small.
 * Rows are in the reference's push order when run serially. */
ORC_API int64_t orc_build_neighbour_list(const orc_box *b, const orc_table *t, int64_t n,
                                         const double *pos, const int32_t *types, double extra,
                                         int64_t *nbr_start, int32_t *nbr, int64_t cap) {
    double max_rcut = orc_max_rcut(t) + extra;
    uint64_t nc[3];
    orc_divide_into_cells(b, max_rcut, nc);
    int64_t nx = (int64_t)nc[0], ny = (int64_t)nc[1], nz = (int64_t)nc[2];
    int64_t ncell_total = nx * ny * nz;
    int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cell_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncell_total + 1));
    int64_t *cell_atoms = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_rcut_cells(b, pos, n, nc[0], nc[1], nc[2], cell_of, cell_start, cell_atoms);

    /* two passes (count, fill) over the same traversal so the CSR rows keep push order */
    int64_t *count = (int64_t *)calloc((size_t)(n + 1), sizeof(int64_t));
    int64_t total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            nbr_start[0] = 0;
            for (int64_t i = 0; i < n; ++i) nbr_start[i + 1] = nbr_start[i] + count[i];
            total = nbr_start[n];
            if (total > cap) {
                total = -1;
                break;
            }
            memset(count, 0, sizeof(int64_t) * (size_t)(n + 1));
        }
        for (int64_t cx_i = 0; cx_i < nx; ++cx_i)
            for (int64_t cy_i = 0; cy_i < ny; ++cy_i)
                for (int64_t cz_i = 0; cz_i < nz; ++cz_i) {
                    int64_t current_cell = (cz_i * ny + cy_i) * nx + cx_i;
                    int64_t seen[9];
                    int n_seen;
                    for (int dx = -1; dx <= 1; ++dx) {
                        n_seen = 0; /* tnc[tid].clear() sits inside the dx loop (:410) */
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dz = -1; dz <= 1; ++dz) {
                                int64_t cx_j = rem_euclid(cx_i + dx, nx);
                                int64_t cy_j = rem_euclid(cy_i + dy, ny);
                                int64_t cz_j = rem_euclid(cz_i + dz, nz);
                                int64_t a_neighbour_cell = (cz_j * ny + cy_j) * nx + cx_j;
                                int dup = 0;
                                for (int s = 0; s < n_seen; ++s)
                                    if (seen[s] == a_neighbour_cell) dup = 1;
                                if (dup) continue;
                                seen[n_seen++] = a_neighbour_cell;
                                for (int64_t ii = cell_start[current_cell]; ii < cell_start[current_cell + 1]; ++ii) {
                                    int64_t i = cell_atoms[ii];
                                    for (int64_t jj = cell_start[a_neighbour_cell];
                                         jj < cell_start[a_neighbour_cell + 1]; ++jj) {
                                        int64_t j = cell_atoms[jj];
                                        if (current_cell == a_neighbour_cell && i == j) continue;
                                        double rij[3] = {pos[3 * j] - pos[3 * i],
                                                         pos[3 * j + 1] - pos[3 * i + 1],
                                                         pos[3 * j + 2] - pos[3 * i + 2]};
                                        orc_min_image(b, rij);
                                        int k = table_lookup(t, types[i], types[j]);
                                        if (k < 0) continue;
                                        if (sqrt(norm_squared3(rij)) > t->rcut[k] + extra) continue;
                                        if (pass == 1) nbr[nbr_start[i] + count[i]] = (int32_t)j;
                                        count[i]++;
                                    }
                                }
                            }
                    }
                }
    }
    free(count);
    free(cell_of);
    free(cell_start);
    free(cell_atoms);
    return total;
}

/* All-core variant of orc_build_neighbour_list: the traversal of the reference's rayon closure
 * (ref: lennard_jones.rs:362-412, `(0..n_cells).into_par_iter()`), cells distributed over threads.
 * A row i is only ever written by the thread that owns i's cell, in the same (dx, dy, dz, jj)
 * order as the serial traversal, so the CSR rows are identical to the serial function's, entry
 * for entry.  ONE traversal: rows go into a padded scratch of cap/n entries per atom first and are
 * compacted to CSR afterwards (returns -1 if a row or the total does not fit; the caller retries).
 * Used by the parity tests at sizes where the serial two-pass build takes minutes (4M atoms). */
ORC_API int64_t orc_build_neighbour_list_omp(const orc_box *b, const orc_table *t, int64_t n,
                                             const double *pos, const int32_t *types, double extra,
                                             int64_t *nbr_start, int32_t *nbr, int64_t cap, int n_threads) {
    double max_rcut = orc_max_rcut(t) + extra;
    uint64_t nc[3];
    orc_divide_into_cells(b, max_rcut, nc);
    int64_t nx = (int64_t)nc[0], ny = (int64_t)nc[1], nz = (int64_t)nc[2];
    int64_t ncell_total = nx * ny * nz;
    int64_t width = n > 0 ? cap / n : 0;
    if (width < 1) return -1;
    int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *cell_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncell_total + 1));
    int64_t *cell_atoms = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_rcut_cells(b, pos, n, nc[0], nc[1], nc[2], cell_of, cell_start, cell_atoms);
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    int64_t *count = (int64_t *)calloc((size_t)(n + 1), sizeof(int64_t));
    int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * (size_t)width * (size_t)(n > 0 ? n : 1));
    int overflow = 0;
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads) reduction(| : overflow)
    for (int64_t current_cell = 0; current_cell < ncell_total; ++current_cell) {
        int64_t cx_i = current_cell % nx, cy_i = (current_cell / nx) % ny, cz_i = current_cell / (nx * ny);
        int64_t seen[9];
        int n_seen;
        for (int dx = -1; dx <= 1; ++dx) {
            n_seen = 0; /* tnc[tid].clear() sits inside the dx loop (:410) */
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    int64_t cx_j = rem_euclid(cx_i + dx, nx);
                    int64_t cy_j = rem_euclid(cy_i + dy, ny);
                    int64_t cz_j = rem_euclid(cz_i + dz, nz);
                    int64_t a_neighbour_cell = (cz_j * ny + cy_j) * nx + cx_j;
                    int dup = 0;
                    for (int s = 0; s < n_seen; ++s)
                        if (seen[s] == a_neighbour_cell) dup = 1;
                    if (dup) continue;
                    seen[n_seen++] = a_neighbour_cell;
                    for (int64_t ii = cell_start[current_cell]; ii < cell_start[current_cell + 1]; ++ii) {
                        int64_t i = cell_atoms[ii];
                        for (int64_t jj = cell_start[a_neighbour_cell]; jj < cell_start[a_neighbour_cell + 1]; ++jj) {
                            int64_t j = cell_atoms[jj];
                            if (current_cell == a_neighbour_cell && i == j) continue;
                            double rij[3] = {pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1],
                                             pos[3 * j + 2] - pos[3 * i + 2]};
                            orc_min_image(b, rij);
                            int k = table_lookup(t, types[i], types[j]);
                            if (k < 0) continue;
                            if (sqrt(norm_squared3(rij)) > t->rcut[k] + extra) continue;
                            if (count[i] < width) rows[i * width + count[i]] = (int32_t)j;
                            else overflow = 1;
                            count[i]++;
                        }
                    }
                }
        }
    }
    int64_t total = -1;
    if (!overflow) {
        nbr_start[0] = 0;
        for (int64_t i = 0; i < n; ++i) nbr_start[i + 1] = nbr_start[i] + count[i];
        total = nbr_start[n];
#pragma omp parallel for schedule(static) num_threads(n_threads)
        for (int64_t i = 0; i < n; ++i) memcpy(nbr + nbr_start[i], rows + i * width, sizeof(int32_t) * (size_t)count[i]);
    }
    free(rows);
    free(count);
    free(cell_of);
    free(cell_start);
    free(cell_atoms);
    return total;
}

/* LJVPBuildListManager::compute_potential's list consumer: F_i += -f_ij over the full list,
 * PE = sum(u)/2.  ref: src/potentials/lennard_jones.rs:419-455.  Here the list may have been
 * built with extra = skin, so the reference cutoff test is re-applied per listed pair. */
ORC_API double orc_compute_potential_list(const orc_box *b, const orc_table *t, int64_t n,
                                          const double *pos, const int32_t *types,
                                          const int64_t *nbr_start, const int32_t *nbr,
                                          double *forces) {
    double pe_sum = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double tpe = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
        for (int64_t q = nbr_start[i]; q < nbr_start[i + 1]; ++q) {
            int64_t j = nbr[q];
            double rij[3] = {pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1],
                             pos[3 * j + 2] - pos[3 * i + 2]};
            orc_min_image(b, rij);
            int k = table_lookup(t, types[i], types[j]);
            if (k < 0) continue;
            if (sqrt(norm_squared3(rij)) > t->rcut[k]) continue;
            double u, f[3];
            orc_lj_pair(t->eps[k], t->sigma[k], t->rcut[k], t->shift, rij, &u, f);
            tpe += u;
            fx += -f[0];
            fy += -f[1];
            fz += -f[2];
        }
        forces[3 * i] += fx;
        forces[3 * i + 1] += fy;
        forces[3 * i + 2] += fz;
        pe_sum += tpe;
    }
    return pe_sum / 2.0;
}

/* Atoms::kinetic_energy.  ref: src/atoms/properties.rs:17-24 */
ORC_API double orc_kinetic_energy(int64_t n, const double *vel, const int32_t *types,
                                  const double *masses) {
    double ek = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double mass = masses[types[i] - 1];
        ek += 0.5 * mass * norm_squared3(&vel[3 * i]);
    }
    return ek;
}

/* Atoms::temerature; dof = 3N; KB_KJPERMOLEKELVIN.  ref: properties.rs:28-30,41-43; constants.rs:3 */
ORC_API double orc_temperature(int64_t n, double kinetic_energy) {
    return (2.0 * kinetic_energy) / ((double)(3 * n) * 0.0083144621);
}

/* virial_tensor().trace() = tr(X * F^T).  ref: src/atoms/properties.rs:49-51,62 */
ORC_API double orc_virial_trace(int64_t n, const double *pos, const double *forces) {
    double t[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < 3; ++d) {
        double acc = 0.0;
        for (int64_t k = 0; k < n; ++k) {
            double p = pos[3 * k + d] * forces[3 * k + d];
            acc = (k == 0) ? p : p + acc;
        }
        t[d] = acc;
    }
    return (t[0] + t[1]) + t[2];
}

/* Atoms::pressure.  ref: src/atoms/properties.rs:61-65 */
ORC_API double orc_pressure(const orc_box *b, int64_t n, const double *pos, const double *forces,
                            double kinetic_energy) {
    double virial = orc_virial_trace(n, pos, forces);
    double volume = orc_box_volume(b);
    return (2.0 * kinetic_energy + virial) / (3.0 * volume);
}

/* PotentialManager::verlet_step_nve.  ref: src/potentials/potential.rs:15-33;
 * current_acceleration: src/atoms/properties.rs:32-39.
 * mode: 0 = faithful serial force driver, 1 = OpenMP driver (n_threads). Returns PE(t+dt). */
ORC_API double orc_verlet_step_nve(const orc_box *b, const orc_table *t, int64_t n, double *pos,
                                   double *vel, double *forces, const int32_t *types,
                                   const double *masses, double dt, int mode, int n_threads) {
    double *a_t = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        double m = masses[types[i] - 1];
        for (int d = 0; d < 3; ++d) a_t[3 * i + d] = 0.0 + forces[3 * i + d] / m;
    }
    double dt2 = dt * dt; /* dt.powi(2) */
    for (int64_t k = 0; k < 3 * n; ++k) pos[k] += (vel[k] * dt) + ((a_t[k] * 0.5) * dt2);
    for (int64_t i = 0; i < n; ++i) orc_wrap_pos(b, &pos[3 * i]);
    memset(forces, 0, sizeof(double) * 3 * (size_t)n);
    double potential_energy = mode == 0 ? orc_compute_potential(b, t, n, pos, types, forces)
                                        : orc_compute_potential_omp(b, t, n, pos, types, forces, n_threads);
    for (int64_t i = 0; i < n; ++i) {
        double m = masses[types[i] - 1];
        for (int d = 0; d < 3; ++d) {
            double a_tdt = 0.0 + forces[3 * i + d] / m;
            vel[3 * i + d] += ((a_t[3 * i + d] + a_tdt) * 0.5) * dt;
        }
    }
    free(a_t);
    return potential_energy;
}

/* Simulation::run, NVE arm: step-0 compute_potential on the input positions (not wrapped),
 * then `steps` x (verlet_step_nve; KE; T; P).  ref: src/simulation.rs:8-88.
 * thermo[(s)*5 + {0:PE,1:KE,2:H,3:T,4:P}] for s = 1..steps at row s; row 0 holds step-0 PE only. */
ORC_API void orc_run_nve(const orc_box *b, const orc_table *t, int64_t n, double *pos, double *vel,
                         double *forces, const int32_t *types, const double *masses, double dt,
                         int64_t steps, int mode, int n_threads, double *thermo) {
    double pe0 = mode == 0 ? orc_compute_potential(b, t, n, pos, types, forces)
                           : orc_compute_potential_omp(b, t, n, pos, types, forces, n_threads);
    thermo[0] = pe0;
    thermo[1] = thermo[2] = thermo[3] = thermo[4] = 0.0;
    for (int64_t s = 0; s < steps; ++s) {
        double pe = orc_verlet_step_nve(b, t, n, pos, vel, forces, types, masses, dt, mode, n_threads);
        double ke = orc_kinetic_energy(n, vel, types, masses);
        double *row = &thermo[(s + 1) * 5];
        row[0] = pe;
        row[1] = ke;
        row[2] = pe + ke;
        row[3] = orc_temperature(n, ke);
        row[4] = orc_pressure(b, n, pos, forces, ke);
    }
}

/* ---- NVT: Nose-Hoover chain (src/ensemble/nvt.rs) and verlet_step_nvt_nhc (src/potentials/potential.rs:35-58) ---- */
typedef struct {
    int32_t chain_size; /* always 3: nvt.rs:108 */
    int32_t pad;
    double start_temperature, end_temperature, target_temperature;
    double xi[3], eta[3], g[3], q[3];
} orc_nhc;

static const double ORC_KB = 0.0083144621;

/* NHThermostatChain::new via new_from_args: target = start, chain_size = 3.  ref: nvt.rs:21-56,99-112 */
ORC_API void orc_nhc_new(orc_nhc *c, double start_temperature, double end_temperature, double tau) {
    memset(c, 0, sizeof *c);
    c->chain_size = 3;
    c->start_temperature = start_temperature;
    c->end_temperature = end_temperature;
    c->target_temperature = start_temperature;
    double q_value = ORC_KB * c->target_temperature * (tau * tau);
    double p10 = 1.0;
    for (int i = 0; i < 3; ++i) {
        c->q[i] = q_value / p10; /* (10.0).powi(i): 1, 10, 100 exactly */
        p10 *= 10.0;
    }
}

/* ref: nvt.rs:59-69 */
ORC_API void orc_nhc_compute_forces(orc_nhc *c, double kinetic_energy, int64_t n_atoms) {
    c->g[0] = 2.0 * kinetic_energy - ((double)(n_atoms * 3)) * ORC_KB * c->target_temperature;
    for (int j = 1; j < c->chain_size; ++j)
        c->g[j] = c->q[j - 1] * (c->xi[j - 1] * c->xi[j - 1]) - ORC_KB * c->target_temperature;
}

/* ref: nvt.rs:72-80 -- xi is ASSIGNED (not incremented) and eta never advances: reproduced as written */
ORC_API void orc_nhc_propagate_half_step(orc_nhc *c, double timestep) {
    int j = c->chain_size - 1;
    c->xi[j] = 0.5 * timestep * c->g[j] / c->q[j];
    for (int l = j - 1; l >= 0; --l)
        c->xi[l] = (0.5 * timestep * c->g[l] / c->q[l]) * exp(-0.25 * timestep * c->xi[l + 1]);
}

/* ref: nvt.rs:82-97 */
ORC_API double orc_nhc_kinetic_energy(const orc_nhc *c) {
    double ke = 0.0;
    for (int i = 0; i < c->chain_size; ++i) ke += 0.5 * c->q[i] * (c->xi[i] * c->xi[i]);
    return ke;
}
ORC_API double orc_nhc_potential_energy(const orc_nhc *c, int64_t n_atoms) {
    double pe = (double)(n_atoms * 3) * ORC_KB * c->target_temperature * c->eta[0];
    for (int i = 1; i < c->chain_size; ++i) pe += ORC_KB * c->target_temperature * c->eta[i];
    return pe;
}

/* ref: nvt.rs:114-123 */
ORC_API void orc_nhc_calculate_target_temperature(orc_nhc *c, int64_t current_timestep, int64_t total_timesteps) {
    c->target_temperature = c->start_temperature +
                            ((c->end_temperature - c->start_temperature) / (double)total_timesteps) * (double)current_timestep;
}

/* PotentialManager::verlet_step_nvt_nhc.  ref: src/potentials/potential.rs:35-58 */
ORC_API double orc_verlet_step_nvt_nhc(const orc_box *b, const orc_table *t, int64_t n, double *pos, double *vel,
                                       double *forces, const int32_t *types, const double *masses, double dt,
                                       orc_nhc *nhc, int mode, int n_threads) {
    double kinetic_energy = orc_kinetic_energy(n, vel, types, masses);
    orc_nhc_compute_forces(nhc, kinetic_energy, n);
    orc_nhc_propagate_half_step(nhc, dt);
    double scale = exp(-0.5 * dt * nhc->xi[0]);
    for (int64_t k = 0; k < 3 * n; ++k) vel[k] = vel[k] * scale;
    double potential_energy = orc_verlet_step_nve(b, t, n, pos, vel, forces, types, masses, dt, mode, n_threads);
    for (int64_t k = 0; k < 3 * n; ++k) vel[k] = vel[k] * scale;
    kinetic_energy = orc_kinetic_energy(n, vel, types, masses);
    orc_nhc_compute_forces(nhc, kinetic_energy, n);
    orc_nhc_propagate_half_step(nhc, dt);
    return potential_energy;
}

/* Simulation::run, NVT arm (src/simulation.rs:8-115): rows as orc_run_nve with H = PE + KE + nhc KE + nhc PE. */
ORC_API void orc_run_nvt(const orc_box *b, const orc_table *t, int64_t n, double *pos, double *vel, double *forces,
                         const int32_t *types, const double *masses, double dt, int64_t steps, orc_nhc *nhc, int mode,
                         int n_threads, double *thermo) {
    double pe0 = mode == 0 ? orc_compute_potential(b, t, n, pos, types, forces)
                           : orc_compute_potential_omp(b, t, n, pos, types, forces, n_threads);
    thermo[0] = pe0;
    thermo[1] = thermo[2] = thermo[3] = thermo[4] = 0.0;
    for (int64_t s = 0; s < steps; ++s) {
        double pe = orc_verlet_step_nvt_nhc(b, t, n, pos, vel, forces, types, masses, dt, nhc, mode, n_threads);
        orc_nhc_calculate_target_temperature(nhc, s, steps); /* simulation.rs:55 */
        double ke = orc_kinetic_energy(n, vel, types, masses);
        double *row = &thermo[(s + 1) * 5];
        row[0] = pe;
        row[1] = ke;
        row[2] = pe + ke + orc_nhc_kinetic_energy(nhc) + orc_nhc_potential_energy(nhc, n);
        row[3] = orc_temperature(n, ke);
        row[4] = orc_pressure(b, n, pos, forces, ke);
    }
}

/* ---- NPT: MTK barostat (src/ensemble/npt.rs), scale_box (src/atoms/transformations.rs:6-15),
 * pressure_tensor (src/atoms/properties.rs:45-59), verlet_step_npt_mtk (src/potentials/potential.rs:112-135).
 *
 * Third-party arithmetic restated here (source not under /root/reference): nalgebra 0.34.1 `Matrix3::exp`
 * (linalg/exp.rs), which follows Al-Mohy & Higham, "A New Scaling and Squaring Algorithm for the Matrix
 * Exponential" (SIAM J. Matrix Anal. Appl. 31, 2009) in the arrangement of scipy.linalg.expm: Pade order 3/5/7/9
 * chosen from d4 = |A^4|_1^(1/4), d6 = |A^6|_1^(1/6), d8, d10 against the theta_m thresholds and ell(A, m) == 0,
 * else order 13 with 2^s scaling and squaring; (V - U) X = (V + U) solved by LU with partial pivoting.
 * The barostat arguments here are O(1e-6 .. 1e-2), i.e. always the order-3 branch.  PARITY UNPINNED like the rest. */
typedef struct {
    double target_pressure[9]; /* column-major like nalgebra */
    double momentum[9];
    double w;
} orc_mtk;

/* C = A * B for column-major 3x3: per output column a gemv in column-axpy order (nalgebra static gemm) */
static void matmul3(const double *a, const double *b, double *c) {
    double t[9];
    for (int j = 0; j < 3; ++j) matvec3(a, &b[3 * j], &t[3 * j]);
    memcpy(c, t, sizeof t);
}
static double onenorm3(const double *a) {
    double best = 0.0;
    for (int j = 0; j < 3; ++j) {
        double s = (fabs(a[3 * j]) + fabs(a[3 * j + 1])) + fabs(a[3 * j + 2]);
        if (s > best) best = s;
    }
    return best;
}
/* ell(A, m) of Al-Mohy & Higham (2009), eq. (3.11)-(3.12), with the exact 1-norm of |A|^(2m+1) */
static int expm_ell(const double *a, int m) {
    /* C(2p, p) * (2p + 1)! for p = 2m + 1 */
    static const double abs_c_recip[14] = {0, 0, 0, 4487938430976000.0, 0, 1.8236839872145106e+28, 0, 1.275506339396217e+42, 0,
                                           7.209685231212166e+56, 0, 0, 0, 2.4719128253168207e+88};
    double aa[9], pw[9];
    for (int k = 0; k < 9; ++k) aa[k] = fabs(a[k]);
    memcpy(pw, aa, sizeof pw);
    for (int k = 1; k < 2 * m + 1; ++k) matmul3(pw, aa, pw);
    double a1 = onenorm3(a);
    if (a1 == 0.0) return 0;
    double alpha = onenorm3(pw) / (a1 * abs_c_recip[m]);
    if (alpha == 0.0) return 0;
    double v = ceil(log2(alpha / ldexp(1.0, -53)) / (2.0 * m));
    return v > 0.0 ? (int)v : 0;
}
/* X = Q^-1 P, LU with partial (row) pivoting, column by column */
static void solve_lu3(const double *q_in, const double *p, double *x) {
    double lu[9];
    int perm[3] = {0, 1, 2};
    memcpy(lu, q_in, sizeof lu);
#define LU(r, c) lu[(c) * 3 + (r)]
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int r = k + 1; r < 3; ++r)
            if (fabs(LU(r, k)) > fabs(LU(piv, k))) piv = r;
        if (piv != k) {
            for (int c = 0; c < 3; ++c) {
                double t = LU(k, c);
                LU(k, c) = LU(piv, c);
                LU(piv, c) = t;
            }
            int t = perm[k];
            perm[k] = perm[piv];
            perm[piv] = t;
        }
        for (int r = k + 1; r < 3; ++r) {
            LU(r, k) = LU(r, k) / LU(k, k);
            for (int c = k + 1; c < 3; ++c) LU(r, c) -= LU(r, k) * LU(k, c);
        }
    }
    for (int j = 0; j < 3; ++j) {
        double y[3];
        for (int r = 0; r < 3; ++r) y[r] = p[3 * j + perm[r]];
        for (int r = 1; r < 3; ++r)
            for (int c = 0; c < r; ++c) y[r] -= LU(r, c) * y[c];
        for (int r = 2; r >= 0; --r) {
            for (int c = r + 1; c < 3; ++c) y[r] -= LU(r, c) * y[c];
            y[r] = y[r] / LU(r, r);
        }
        for (int r = 0; r < 3; ++r) x[3 * j + r] = y[r];
    }
#undef LU
}
/* U = A * (sum_k b[2k+1] A^(2k)), V = sum_k b[2k] A^(2k) for the Pade order m in {3,5,7,9} */
static void expm_pade_uv(const double *a, int m, const double *b, double *u, double *v) {
    double a2[9], pw[9], eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, su[9], sv[9];
    matmul3(a, a, a2);
    for (int k = 0; k < 9; ++k) {
        su[k] = eye[k] * b[1];
        sv[k] = eye[k] * b[0];
    }
    memcpy(pw, eye, sizeof pw);
    for (int k = 1; 2 * k <= m; ++k) {
        matmul3(pw, a2, pw);
        for (int e = 0; e < 9; ++e) {
            su[e] = pw[e] * b[2 * k + 1] + su[e];
            sv[e] = pw[e] * b[2 * k] + sv[e];
        }
    }
    matmul3(a, su, u);
    memcpy(v, sv, sizeof sv);
}
ORC_API void orc_mat3_exp(const double *a, double *out) {
    static const double b3[] = {120., 60., 12., 1.};
    static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
    static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                                 1323241920.,        40840800.,          960960.,           16380., 182., 1.};
    double a2[9], a4[9], a6[9], a8[9], a10[9], u[9], v[9], p[9], q[9];
    matmul3(a, a, a2);
    matmul3(a2, a2, a4);
    matmul3(a4, a2, a6);
    const double d4 = pow(onenorm3(a4), 0.25), d6 = pow(onenorm3(a6), 1.0 / 6.0);
    int order = 0;
    const double eta1 = fmax(d4, d6);
    if (eta1 < 1.495585217958292e-2 && expm_ell(a, 3) == 0) order = 3;
    else if (eta1 < 2.539398330063230e-1 && expm_ell(a, 5) == 0) order = 5;
    else {
        matmul3(a4, a4, a8);
        const double d8 = pow(onenorm3(a8), 0.125);
        const double eta3 = fmax(d6, d8);
        if (eta3 < 9.504178996162932e-1 && expm_ell(a, 7) == 0) order = 7;
        else if (eta3 < 2.097847961257068 && expm_ell(a, 9) == 0) order = 9;
        else {
            matmul3(a4, a6, a10);
            const double d10 = pow(onenorm3(a10), 0.1);
            const double eta4 = fmax(d8, d10), eta5 = fmin(eta3, eta4);
            const double theta13 = 4.25;
            int s = 0;
            if (eta5 > 0.0) {
                double v2 = ceil(log2(eta5 / theta13));
                if (v2 > 0.0) s = (int)v2;
            }
            double as[9];
            for (int k = 0; k < 9; ++k) as[k] = ldexp(a[k], -s);
            s += expm_ell(as, 13);
            for (int k = 0; k < 9; ++k) as[k] = ldexp(a[k], -s);
            double b2[9], b4[9], b6[9], eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t1[9], t2[9];
            matmul3(as, as, b2);
            matmul3(b2, b2, b4);
            matmul3(b4, b2, b6);
            for (int k = 0; k < 9; ++k) t1[k] = b13[13] * b6[k] + b13[11] * b4[k] + b13[9] * b2[k];
            matmul3(b6, t1, t2);
            for (int k = 0; k < 9; ++k) t2[k] = t2[k] + b13[7] * b6[k] + b13[5] * b4[k] + b13[3] * b2[k] + b13[1] * eye[k];
            matmul3(as, t2, u);
            for (int k = 0; k < 9; ++k) t1[k] = b13[12] * b6[k] + b13[10] * b4[k] + b13[8] * b2[k];
            matmul3(b6, t1, t2);
            for (int k = 0; k < 9; ++k) v[k] = t2[k] + b13[6] * b6[k] + b13[4] * b4[k] + b13[2] * b2[k] + b13[0] * eye[k];
            for (int k = 0; k < 9; ++k) {
                p[k] = u[k] + v[k];
                q[k] = v[k] - u[k];
            }
            solve_lu3(q, p, out);
            for (int k = 0; k < s; ++k) matmul3(out, out, out);
            return;
        }
    }
    expm_pade_uv(a, order, order == 3 ? b3 : order == 5 ? b5 : order == 7 ? b7 : b9, u, v);
    for (int k = 0; k < 9; ++k) {
        p[k] = u[k] + v[k];
        q[k] = v[k] - u[k];
    }
    solve_lu3(q, p, out);
}

/* math::symmetrize.  ref: src/math.rs:41-43 */
static void symmetrize3(const double *a, double *out) {
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[c * 3 + r] = (a[c * 3 + r] + a[r * 3 + c]) * 0.5;
    memcpy(out, t, sizeof t);
}

/* A(3xN) * B(3xN)^T, nalgebra's generic gemm: per output column c a gemv over atoms, out[r,c] = sum_k A[r,k] B[c,k]
 * accumulated sequentially in k.  ref: kinetic_tensor / virial_tensor, src/atoms/properties.rs:45-51 */
static void outer_sum3(int64_t n, const double *a, const double *b, double *out) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            double acc = 0.0;
            for (int64_t k = 0; k < n; ++k) {
                double p = a[3 * k + r] * b[3 * k + c];
                acc = (k == 0) ? p : p + acc;
            }
            out[c * 3 + r] = acc;
        }
}

/* Atoms::pressure_tensor (the kinetic tensor carries no mass, as in the reference).  ref: properties.rs:53-59 */
ORC_API void orc_pressure_tensor(const orc_box *b, int64_t n, const double *pos, const double *vel, const double *forces,
                                 double *out) {
    double kt[9], vt[9];
    outer_sum3(n, vel, vel, kt);
    outer_sum3(n, pos, forces, vt);
    double volume = orc_box_volume(b);
    for (int k = 0; k < 9; ++k) out[k] = (kt[k] + vt[k]) / volume;
}

/* MTKBarostat::new via new_from_args (target temperature = the thermostat's start temperature).  ref: npt.rs:24-43,67-88 */
ORC_API void orc_mtk_new(orc_mtk *m, const double *target_pressure9, double tau, int64_t n_atoms, double target_temp) {
    memcpy(m->target_pressure, target_pressure9, sizeof m->target_pressure);
    memset(m->momentum, 0, sizeof m->momentum);
    m->w = ((double)(3 * n_atoms)) * ORC_KB * target_temp * (tau * tau);
}

/* ref: npt.rs:45-50 */
ORC_API void orc_mtk_delta_momentum(const orc_mtk *m, const orc_box *b, int64_t n, const double *pos, const double *vel,
                                    const double *forces, double dt, double *out) {
    double p[9];
    orc_pressure_tensor(b, n, pos, vel, forces, p);
    double f = orc_box_volume(b) * 0.5 * dt;
    for (int k = 0; k < 9; ++k) p[k] = (p[k] - m->target_pressure[k]) * f;
    symmetrize3(p, out);
}

/* ref: npt.rs:52-58 */
ORC_API void orc_mtk_scale(const orc_mtk *m, double dt, int velocity_scaling, double *out) {
    double e[9];
    for (int k = 0; k < 9; ++k) e[k] = m->momentum[k] / m->w;
    symmetrize3(e, e);
    double factor = velocity_scaling ? -0.5 : 1.0;
    for (int k = 0; k < 9; ++k) e[k] = e[k] * factor * dt;
    orc_mat3_exp(e, out);
}

/* ref: npt.rs:60-66.  tr(M M^T) and tr(P^T h), each product element in gemv order */
ORC_API double orc_mtk_kinetic_energy(const orc_mtk *m) {
    double mt[9], pr[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) mt[c * 3 + r] = m->momentum[r * 3 + c];
    matmul3(m->momentum, mt, pr);
    return ((pr[0] + pr[4]) + pr[8]) / (2.0 * m->w);
}
ORC_API double orc_mtk_potential_energy(const orc_mtk *m, const double *h9) {
    double pt[9], pr[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) pt[c * 3 + r] = m->target_pressure[r * 3 + c];
    matmul3(pt, h9, pr);
    return (pr[0] + pr[4]) + pr[8];
}

/* Atoms::scale_box.  ref: src/atoms/transformations.rs:6-15.  Returns 1 if the new box is singular. */
ORC_API int orc_scale_box(orc_box *b, const double *scale9, int64_t n, double *pos) {
    double hn[9];
    matmul3(scale9, b->h, hn);
    orc_box nb;
    if (orc_box_new(hn, b->pbc, &nb)) return 1;
    for (int64_t i = 0; i < n; ++i) {
        double s[3];
        matvec3(b->hinv, &pos[3 * i], s);
        matvec3(nb.h, s, &pos[3 * i]);
    }
    *b = nb;
    return 0;
}

/* PotentialManager::verlet_step_npt_mtk.  ref: src/potentials/potential.rs:112-135 */
ORC_API double orc_verlet_step_npt_mtk(orc_box *b, const orc_table *t, int64_t n, double *pos, double *vel, double *forces,
                                       const int32_t *types, const double *masses, double dt, orc_mtk *mtk, orc_nhc *nhc,
                                       int mode, int n_threads) {
    double dm[9], scale[9], scale_h[9], tv[3];
    orc_mtk_delta_momentum(mtk, b, n, pos, vel, forces, dt, dm);
    for (int k = 0; k < 9; ++k) mtk->momentum[k] += dm[k];
    orc_mtk_scale(mtk, dt, 1, scale);
    for (int64_t i = 0; i < n; ++i) {
        matvec3(scale, &vel[3 * i], tv);
        memcpy(&vel[3 * i], tv, sizeof tv);
    }
    orc_mtk_scale(mtk, dt, 0, scale_h);
    orc_scale_box(b, scale_h, n, pos);
    double potential_energy = orc_verlet_step_nvt_nhc(b, t, n, pos, vel, forces, types, masses, dt, nhc, mode, n_threads);
    for (int64_t i = 0; i < n; ++i) {
        matvec3(scale, &vel[3 * i], tv);
        memcpy(&vel[3 * i], tv, sizeof tv);
    }
    orc_mtk_delta_momentum(mtk, b, n, pos, vel, forces, dt, dm);
    for (int k = 0; k < 9; ++k) mtk->momentum[k] += dm[k];
    return potential_energy;
}

/* Simulation::run, NPT arm (src/simulation.rs:8-115): rows as orc_run_nve with
 * H = PE + KE + nhc KE + nhc PE + mtk KE + mtk PE; h_trace (may be NULL) receives the box after every step (9 per row). */
ORC_API void orc_run_npt(orc_box *b, const orc_table *t, int64_t n, double *pos, double *vel, double *forces,
                         const int32_t *types, const double *masses, double dt, int64_t steps, orc_mtk *mtk, orc_nhc *nhc,
                         int mode, int n_threads, double *thermo, double *h_trace) {
    double pe0 = mode == 0 ? orc_compute_potential(b, t, n, pos, types, forces)
                           : orc_compute_potential_omp(b, t, n, pos, types, forces, n_threads);
    thermo[0] = pe0;
    thermo[1] = thermo[2] = thermo[3] = thermo[4] = 0.0;
    if (h_trace) memcpy(h_trace, b->h, sizeof(double) * 9);
    for (int64_t s = 0; s < steps; ++s) {
        double pe = orc_verlet_step_npt_mtk(b, t, n, pos, vel, forces, types, masses, dt, mtk, nhc, mode, n_threads);
        orc_nhc_calculate_target_temperature(nhc, s, steps); /* simulation.rs:60 */
        double ke = orc_kinetic_energy(n, vel, types, masses);
        double *row = &thermo[(s + 1) * 5];
        row[0] = pe;
        row[1] = ke;
        row[2] = pe + ke + orc_nhc_kinetic_energy(nhc) + orc_nhc_potential_energy(nhc, n) + orc_mtk_kinetic_energy(mtk) +
                 orc_mtk_potential_energy(mtk, b->h);
        row[3] = orc_temperature(n, ke);
        row[4] = orc_pressure(b, n, pos, forces, ke);
        if (h_trace) memcpy(&h_trace[(s + 1) * 9], b->h, sizeof(double) * 9);
    }
}

/* Largest double T with sqrt(T) <= rc (correctly-rounded sqrt is monotone), so that
 * `sqrt(r2) > rc`  <=>  `r2 > T` exactly.  Test helper for the sqrt-free device predicate. */
ORC_API double orc_rcut_threshold(double rc) {
    double t = rc * rc;
    while (sqrt(t) > rc) t = nextafter(t, 0.0);
    while (sqrt(nextafter(t, INFINITY)) <= rc) t = nextafter(t, INFINITY);
    return t;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
