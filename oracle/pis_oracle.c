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
