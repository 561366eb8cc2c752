/*
 * pisb200.h -- C ABI of libpisb200.so: the B200-native (sm_100a) implementation of the PIS
 * per-timestep MD hot path (cell binning -> Verlet neighbour list -> Lennard-Jones
 * force/energy/virial -> velocity-Verlet NVE), f64, device-resident.
 *
 * The reference (vivekadishankara/pis, Rust) has no FFI; its hot path sits behind the Rust trait
 * `PotentialManager` (src/potentials/potential.rs:12-136).  Each entry point below names the
 * reference interface it replaces (paths relative to the reference repo).  INTEGRATION.md shows
 * the Rust `extern "C"` block + `impl PotentialManager for LJCudaManager` a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns a PISB_* code (0 = ok);
 * pisb_last_error(h) gives the message for the last failure on that handle (h may be NULL for
 * pisb_create failures).  Host arrays are xyz-interleaved 3N doubles in ORIGINAL atom order,
 * exactly nalgebra's Matrix3xX<f64> storage (src/atoms/new.rs:9-17); types are 1-based.
 * A handle is not re-entrant; independent handles may be used from different threads.
 * There is NO CPU fallback: without a CUDA device every call fails with PISB_ERR_NO_DEVICE.
 */
#ifndef PISB200_H
#define PISB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PISB_API __attribute__((visibility("default")))
#else
#define PISB_API
#endif

typedef struct pisb_handle pisb_t;

enum {
    PISB_OK = 0,
    PISB_ERR_INVALID = 1,   /* bad argument / degenerate input (box smaller than 2 cutoffs, ...) */
    PISB_ERR_CUDA = 2,      /* a CUDA runtime call failed; see pisb_last_error */
    PISB_ERR_NO_DEVICE = 3, /* no usable CUDA device: the product path has no CPU fallback */
    PISB_ERR_CAPACITY = 4,  /* neighbour-list / cell capacity exceeded and regrow failed */
    PISB_ERR_STATE = 5,     /* call order violated (e.g. step before upload / set_box) */
    PISB_ERR_COMM = 6       /* multi-GPU communicator failure */
};

/* Per-step observables.  Replaces what Simulation::output computes on the host every step
 * (src/simulation.rs:79-82): step_potential, Atoms::kinetic_energy (src/atoms/properties.rs:17-24)
 * and the trace of Atoms::virial_tensor (properties.rs:49-51,62).  virial_pair is the periodic
 * pair virial sum_{i<j} r_ij . f_ij (north_star: "pair force/energy/virial"), which the
 * reference does not compute. */
typedef struct {
    double pe;          /* total shifted LJ potential energy (return value of verlet_step_nve) */
    double ke;          /* sum 0.5 m v^2 */
    double virial_ref;  /* tr(X F^T) with wrapped absolute positions, as the reference's pressure uses */
    double virial_pair; /* sum over pairs r_ij . f_ij */
} pisb_thermo;

typedef struct {
    int64_t n_atoms;        /* owned atoms on this handle */
    int64_t n_ghost;        /* ghost atoms (multi-GPU), 0 on a single GPU */
    int64_t n_cells[3];     /* cell grid */
    int64_t list_capacity;  /* neighbour slots per atom (K_max) */
    int64_t max_neighbours; /* longest list row any build has produced since the capacity was last set (a running maximum) */
    int64_t capacity_growths; /* times the capacity was raised AHEAD of an overflow (longest row within 12 % of it) */
    int64_t n_builds;       /* neighbour-list builds so far */
    int64_t n_steps;        /* NVE steps so far */
    int64_t n_launches;     /* kernels launched by this handle so far */
    int64_t device_bytes;   /* device memory held */
    int64_t missing_type_pairs; /* unordered type pairs (i <= j) that atoms of this handle populate but the table does not hold:
                                 * their pairs are skipped.  (The reference prints a line per candidate pair and skips it,
                                 * lennard_jones.rs:216-222; here it is a count, taken when the types are uploaded.) */
} pisb_stats_t;

/* kernel classes for pisb_timings */
enum {
    PISB_K_INTEGRATE = 0, /* fused velocity-Verlet kick/drift + wrap + thermo + displacement check */
    PISB_K_BIN = 1,       /* cell index + histogram */
    PISB_K_SORT = 2,      /* scan + scatter + in-cell order + permute */
    PISB_K_BUILD = 3,     /* neighbour-list build */
    PISB_K_FORCE = 4,     /* LJ force/energy/virial */
    PISB_K_REDUCE = 5,    /* thermo finalisation */
    PISB_K_HALO = 6,      /* halo pack/unpack + exchange (multi-GPU) */
    PISB_K_COPY = 7,      /* host<->device staging conversion */
    PISB_K_COUNT = 8
};

PISB_API const char *pisb_version(void);
PISB_API int pisb_device_count(void);

/* Create a device-backed LJ manager.
 * Replaces: LJVOffsetManager::new + insert((i,j), LennardJones::new(eps, sigma, rcut, shift))
 *   (src/potentials/lennard_jones.rs:14-30,182-184; src/system.rs:134-164;
 *    src/readers/input_file/commands.rs:171,263-292) and Atoms.masses (src/atoms/new.rs:12).
 * eps/sigma/rcut/present are dense n_types x n_types, entry for 1-based pair (i,j) at
 * [(i-1)*n_types + (j-1)], stored AS GIVEN: lookups read (min,max) like the reference's sorted
 * HashMap key (src/potentials/potential.rs:181-192), so an entry given as (2,1) is never found.
 * present[k] == 0 => no potential for that pair: the pair is skipped (reference prints a line and
 * skips, lennard_jones.rs:216-222).  shift: energy shift at rcut (always 1 in the reference).
 * skin: Verlet skin distance (NEW; the reference rebuilds cells every call and has no skin).
 *   skin == 0 reproduces the reference's "rebuild every call" behaviour. */
PISB_API int pisb_create(int device, int n_types, const double *mass, const double *eps,
                         const double *sigma, const double *rcut, const unsigned char *present,
                         int shift, double skin, pisb_t **out);
PISB_API int pisb_destroy(pisb_t *h);
PISB_API const char *pisb_last_error(pisb_t *h);

/* Replaces: SimulationBox{h, h_inv, pbc} (src/simulation_box.rs:5-15).  h9/hinv9 column-major.
 * h_inv is an INPUT (the reference caches nalgebra's try_inverse; the device never recomputes it). */
PISB_API int pisb_set_box(pisb_t *h, const double *h9, const double *hinv9, const int *pbc3);

/* Replaces: the Atoms the trait methods borrow (`&mut Atoms`, src/atoms/new.rs:9-17).
 * pos/vel/force are 3N xyz-interleaved doubles in original order; vel/force may be NULL (= zeros);
 * types 1-based.  Invalidates the neighbour list unless the atom count, box and list are still
 * valid for the new positions (checked on the device against the positions at the last build). */
PISB_API int pisb_upload(pisb_t *h, int64_t n, const double *pos, const double *vel,
                         const double *force, const int32_t *types);

/* Replaces: PotentialManager::compute_potential(&self, &mut Atoms) -> f64
 *   (src/potentials/potential.rs:13; live impl LJVOffsetManager, lennard_jones.rs:186-244).
 * accumulate != 0: forces are ADDED to the device force buffer (the reference semantics: caller
 * zeroes, potential.rs:24); accumulate == 0: forces are overwritten.  Positions are NOT wrapped
 * (the reference's step-0 call runs on the input positions, src/simulation.rs:28). */
PISB_API int pisb_compute(pisb_t *h, int accumulate, double *pe);

/* Replaces: PotentialManager::verlet_step_nve(&self, &mut Atoms, dt) -> f64, nsteps times
 *   (src/potentials/potential.rs:15-33), fused with the per-step observables of
 *   Simulation::output (src/simulation.rs:79-82).  State stays on the device; out (may be NULL)
 *   receives nsteps records; out[s].pe is the value verlet_step_nve returns at step s. */
PISB_API int pisb_step_nve(pisb_t *h, double dt, int64_t nsteps, pisb_thermo *out);

/* Nose-Hoover chain state == NHThermostatChain (src/ensemble/nvt.rs:4-17; chain_size is 3, nvt.rs:108). */
typedef struct {
    int32_t chain_size;
    int32_t pad;
    double start_temperature, end_temperature, target_temperature;
    double xi[3], eta[3], g[3], q[3];
} pisb_nhc;

/* Replaces: NHThermostatChain::new_from_args (nvt.rs:21-56,99-112): target = start, q_i = kB T tau^2 / 10^i. */
PISB_API int pisb_nhc_init(pisb_nhc *out, double start_temperature, double end_temperature, double tau);

/* Replaces: PotentialManager::verlet_step_nvt_nhc (src/potentials/potential.rs:35-58) followed by
 *   NHThermostatChain::calculate_target_temperature(i, total_steps) (src/simulation.rs:53-57), nsteps times,
 *   device-resident (the chain lives on the GPU during the batch).  first_step: index i of the first step
 *   of this batch in the run; total_steps: the run length the temperature ramp refers to.
 *   out[s] as in pisb_step_nve (ke is the kinetic energy after the second thermostat scaling);
 *   nhc_energy[s] (may be NULL) = nhc.kinetic_energy() + nhc.potential_energy(n), the thermostat's share of the
 *   Hamiltonian (simulation.rs:101-104).  On bricks (after pisb_comm_init) the call is collective: every rank advances a
 *   replica of the chain from the all-reduced kinetic energy (one 4-number all-reduce per step) and n is the global atom count;
 *   out[] then holds global sums on every rank. */
PISB_API int pisb_step_nvt_nhc(pisb_t *h, double dt, int64_t nsteps, pisb_nhc *chain, int64_t first_step,
                               int64_t total_steps, pisb_thermo *out, double *nhc_energy);

/* MTK barostat state == MTKBarostat (src/ensemble/npt.rs:11-22); 3x3 matrices column-major like nalgebra. */
typedef struct {
    double target_pressure[9];
    double momentum[9];
    double w;
} pisb_mtk;

/* Replaces: MTKBarostat::new via new_from_args (npt.rs:24-43,67-88): momentum = 0, w = 3 N kB T tau^2 with
 *   T = the thermostat's start temperature.  `fix ... npt ... iso p p tau` gives target_pressure = p * identity
 *   (src/readers/input_file/commands.rs:429-447). */
PISB_API int pisb_mtk_init(pisb_mtk *out, const double *target_pressure9, double tau, int64_t n_atoms,
                           double target_temperature);

/* Replaces: PotentialManager::verlet_step_npt_mtk (src/potentials/potential.rs:112-135) followed by
 *   calculate_target_temperature(i, total_steps) (src/simulation.rs:58-62), nsteps times: barostat half kick from
 *   Atoms::pressure_tensor (properties.rs:45-59), v = exp(-dt/2 eta_dot) v, Atoms::scale_box with exp(dt eta_dot)
 *   (transformations.rs:6-15), verlet_step_nvt_nhc, v scaling and barostat half kick again.  The box held by the
 *   handle CHANGES (and in general becomes triclinic: the pressure tensor has off-diagonal terms); read it back
 *   with pisb_get_box.  out[s].ke / virial_ref are taken AFTER the final velocity scaling, as Simulation::output sees them.
 *   ext_energy[s] (may be NULL) = nhc KE + nhc PE + mtk.kinetic_energy() + mtk.potential_energy(h), the extended
 *   system's share of the Hamiltonian (simulation.rs:105-113); h9_trace (may be NULL) receives the box after every
 *   step (9 doubles per step).  One device->host read of 19 reduced sums per step.  Single-GPU only. */
PISB_API int pisb_step_npt_mtk(pisb_t *h, double dt, int64_t nsteps, pisb_mtk *baro, pisb_nhc *chain,
                               int64_t first_step, int64_t total_steps, pisb_thermo *out, double *ext_energy,
                               double *h9_trace);

/* The box currently held by the handle (either pointer may be NULL): atoms.sim_box.{h, h_inv} after NPT steps. */
PISB_API int pisb_get_box(pisb_t *h, double *h9, double *hinv9);

/* Copy state back in ORIGINAL atom order (any pointer may be NULL).  Replaces reading
 * atoms.positions / velocities / forces on the host (e.g. DumpTraj::write_atoms_info,
 * src/writers/dump_traj.rs:52-66). */
PISB_API int pisb_download(pisb_t *h, double *pos, double *vel, double *force);

/* The same copy, asynchronous: for the dump steps of Simulation::run (src/simulation.rs:31-37,84-86), so that the
 * positions of a dump step travel over PCIe while the next batch of steps already runs.  pisb_download_begin takes a
 * snapshot of the state on the device (in stream order, ORIGINAL atom order) and starts the device->host copies on a
 * second stream; the arrays are valid after pisb_download_end.  One download may be in flight per handle; the
 * destination should be page-locked (pisb_host_register) or the copy degrades to a synchronous one.  Single GPU. */
PISB_API int pisb_download_begin(pisb_t *h, double *pos, double *vel, double *force);
PISB_API int pisb_download_end(pisb_t *h);

/* Page-lock / unlock host arrays the caller owns (nalgebra's Matrix3xX storage of Atoms.positions / velocities /
 * forces, src/atoms/new.rs:9-17) so that the host-buffer step and the downloads run at PCIe line rate and
 * asynchronously.  Thin wrappers over cudaHostRegister / cudaHostUnregister; no handle needed. */
PISB_API int pisb_host_register(void *ptr, size_t bytes);
PISB_API int pisb_host_unregister(void *ptr);

/* Observables of the CURRENT device state (KE, virial_ref from current x, v, F; pe/virial_pair of
 * the last force evaluation).  Replaces Atoms::kinetic_energy / virial_tensor().trace(). */
PISB_API int pisb_thermo_now(pisb_t *h, pisb_thermo *out);

/* `velocity all create <temperature> <seed>` on the device: replaces Atoms::start_velocities
 * (src/atoms/velocities.rs:10-15) for the atoms the handle holds -- Gaussian velocities with sigma_i = sqrt(kB T / m_i)
 * (:17-33), remove_drift (:35-50), rescale_to_temperature (:52-59; kB = 0.0083144621, src/constants.rs:3; 3N degrees of
 * freedom, src/atoms/properties.rs:28-43).  The reference's random stream (rand SmallRng + rand_distr Normal) is third-party
 * and unpinned: the numbers come from this library's generator, keyed by (seed, global atom id, component), so any number of
 * GPUs produces the velocities one GPU would.  Multi-GPU: collective (two small all-reduces).  Forces and list are kept. */
PISB_API int pisb_start_velocities(pisb_t *h, double temperature, uint64_t seed);

/* Strict drop-in for ONE trait call with HOST buffers: upload pos/vel/force (pinned staging),
 * one verlet_step_nve, download pos/vel/force, return PE.  This is what an unmodified
 * Simulation::run (src/simulation.rs:31-37,52) drives through the Rust shim. */
PISB_API int pisb_verlet_step_nve_host(pisb_t *h, int64_t n, double *pos, double *vel,
                                       double *force, const int32_t *types, double dt, double *pe);

/* Test hook: the current Verlet list in ORIGINAL atom ids.  nnbr[n]; nbr[n * cap_per_atom]
 * row-major, rows unsorted.  Semantic source: LJVPBuildListManager::build_neighbour_list
 * (src/potentials/lennard_jones.rs:345-415) with rcut := rcut + skin. */
PISB_API int pisb_neighbours(pisb_t *h, int32_t *nnbr, int32_t *nbr, int64_t cap_per_atom);
/* Force a list rebuild at the next force evaluation (test hook). */
PISB_API int pisb_invalidate_list(pisb_t *h);

PISB_API int pisb_stats(pisb_t *h, pisb_stats_t *out);

/* Per-kernel-class device timing (CUDA events on the handle's stream).  enable: 0/1.
 * pisb_timings drains recorded events: ms[PISB_K_COUNT], launches[PISB_K_COUNT] accumulated since
 * the last pisb_timings_reset. */
PISB_API int pisb_set_profiling(pisb_t *h, int enable);
PISB_API int pisb_timings(pisb_t *h, double *ms, int64_t *launches);
PISB_API int pisb_timings_reset(pisb_t *h);

/* The CUDA stream (cudaStream_t) all work of this handle is launched on, and a full sync. */
PISB_API void *pisb_stream(pisb_t *h);
PISB_API int pisb_synchronize(pisb_t *h);

/* ---- multi-GPU: spatial decomposition, ONE PROCESS PER GPU (new; the reference is single-process) ----
 * The periodic box is cut into grid3[0] x grid3[1] x grid3[2] bricks (each 1 or 2), one per rank;
 * ranks exchange ghost-atom positions every step -- by default with NVLink stores into each peer's CUDA-IPC
 * mapped receive buffer from the packing kernel itself (ncclSend/ncclRecv is the fallback, option halo_mode) --
 * and migrate atoms over NCCL on list-rebuild steps.  Bootstrap: rank 0 calls pisb_comm_unique_id and broadcasts the 128 bytes
 * (torch.distributed / MPI / a file -- plumbing), then every rank calls pisb_comm_init.
 * After pisb_comm_init: pisb_set_box (the GLOBAL box), pisb_upload_owned (this rank's atoms with
 * their global ids; atoms outside the rank's brick are migrated at the first rebuild), then
 * pisb_compute / pisb_step_nve as on one GPU -- thermo records come back globally reduced. */
PISB_API int pisb_comm_unique_id(void *out128, int nbytes);
PISB_API int pisb_comm_init(pisb_t *h, int rank, int nranks, const void *unique_id128, const int *grid3);
PISB_API int pisb_upload_owned(pisb_t *h, int64_t n_own, const double *pos, const double *vel,
                               const double *force, const int32_t *types, const int32_t *global_ids);
/* Owned atoms of this rank (any of pos/vel/force may be NULL), compacted on the device; the order is
 * arbitrary -- global_ids identifies each row.  Arrays hold up to cap atoms; *n_out = count. */
PISB_API int pisb_download_owned(pisb_t *h, int64_t cap, double *pos, double *vel, double *force,
                                 int32_t *global_ids, int64_t *n_out);
/* The same rows, asynchronously: the owned atoms are snapshotted on the device, the copies run on a stream of their own -- in
 * pieces, spread over the steps of the next pisb_step_nve batch (option dump_spread_steps) -- and that batch may start at once;
 * *n_out is valid on return, the arrays after pisb_download_end(h).
 * (Multi-GPU form of pisb_download_begin: a dump frame of Simulation::run, src/simulation.rs:84-86, hidden under the steps
 * that follow.)  Host arrays should be page-locked (pisb_host_register). */
PISB_API int pisb_download_owned_begin(pisb_t *h, int64_t cap, double *pos, double *vel, double *force,
                                       int32_t *global_ids, int64_t *n_out);
/* Test hook: global ids of the owned atoms in device slot order == the row order of pisb_neighbours in
 * multi-GPU mode (where list entries are global ids). */
PISB_API int pisb_owned_ids(pisb_t *h, int64_t cap, int32_t *global_ids, int64_t *n_out);

/* What the current Verlet list holds (built first if need be), counted on the device: out3[0] = listed pairs (sum of the
 * per-atom row lengths, each pair twice: a full list), out3[1] = of those, pairs inside the cutoff at the current
 * positions (`rij.norm() > rcut -> skip`, lennard_jones.rs:224), out3[2] = index words stored (== out3[0] for per-atom
 * rows).  Multi-GPU: this rank's owned atoms. */
PISB_API int pisb_list_stats(pisb_t *h, int64_t *out3);

/* Options (all have working defaults; the kernel selectors exist for A/B measurements and the parity tests):
 *   list_capacity     neighbour slots per atom, 0 = automatic (estimated from the density, grown on overflow)
 *   cuda_graphs       1 (default) = NVE / NVT batches replay CUDA graphs of 8 / 4 / 2 steps, 0 = classic launches
 *   fuse_vv           1 (default) = NVE steps run ONE kernel, the force pass with the velocity-Verlet kick + drift in its
 *                     epilogue (k_force_vv / k_force_q); 0 = force kernel + k_vv
 *   host_pipeline     1 (default) = pisb_verlet_step_nve_host moves x, v, F in chunks and pipelines upload, drift and
 *                     download; 0 = whole-array copies.  host_chunk_atoms = atoms per chunk, 0 = n/8 (>= 65536)
 *   force_variant     0 = automatic (v1 for a triclinic or non-periodic box; otherwise k_force_q up to 75k atoms, one
 *                     thread per atom above), 1 = v1 general all-FP64 in the reference's operation order, 2 = FP32
 *                     pre-filter + queue, 3 = v3 thread per atom, 6 = 8 lanes per atom (whole K-tiles per lane),
 *                     7 = k_force_q: 4 lanes per atom taking one entry of every K-tile each
 *   build_variant     0 = automatic (v3), 1 = v1 general, 2 = scalar FP32 pre-filter, 3 = packed-FP32 pair records
 *   cell_div          cells per list cutoff and dimension: 0 = automatic (2 with build_variant 2/3), 1, 2
 *   dump_spread_steps multi-GPU: steps over which the frame of pisb_download_owned_begin is sent (8; pieces leave after each
 *                     step's decision has reached the host, pisb_download_end sends the rest; 1 = the whole frame at once)
 *   halo_mode         multi-GPU per-step ghost exchange: 0 = peer memory when it can be mapped, 1 = NCCL send/recv,
 *                     2 = peer memory or fail */
PISB_API int pisb_set_option(pisb_t *h, const char *name, double value);

#ifdef __cplusplus
}
#endif
#endif /* PISB200_H */
