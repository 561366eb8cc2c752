// pis_host.hpp -- C++ host side above the C ABI, mirroring the reference's own surface for the hot
// path (the reference is Rust; no rustc in this image, so the host is C++ with the same names,
// argument meaning and error behaviour).  Citations are paths in the reference repository.
//
//   SimulationBox      src/simulation_box.rs
//   Atoms              src/atoms/new.rs, src/atoms/properties.rs (scalar formulas)
//   LennardJones, LJCudaManager  src/potentials/{lennard_jones,potential,kind}.rs  (device-backed)
//   script / data readers   the FORMATS of the reference's input script and LAMMPS data file (behaviour pinned by the cases
//                           of src/tests/command_tests.rs, re-expressed in tests/test_cli.py); own grammar table + tokenizer
//   SimulationContext  src/readers/simulation_context.rs (what a parsed script amounts to)
//   System             src/system.rs (read -> contextualize -> run)
//   Simulation         src/simulation.rs   (NVE, NVT and NPT arms)
//   DumpTraj           src/writers/dump_traj.rs
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pisb200.h"

namespace pis {

constexpr double KB_KJPERMOLEKELVIN = 0.0083144621;  // src/constants.rs:3

// PisError (src/errors.rs:5-100): what() carries the reference's Display text.
struct PisError : std::runtime_error {
    std::string kind;
    PisError(std::string k, const std::string &msg) : std::runtime_error(msg), kind(std::move(k)) {}
};

std::string rust_display_f64(double v);  // Rust `{}` for f64: shortest round-trip, never scientific

struct SimulationBox {
    double h[9];      // column-major like nalgebra
    double h_inv[9];
    bool pbc[3];
    static SimulationBox make(const double h_colmajor[9], const bool pbc[3]);            // SimulationBox::new
    static SimulationBox from_lammps_data(double xlo, double xhi, double ylo, double yhi, double zlo, double zhi,
                                          double xy, double xz, double yz);               // :44-65
    double volume() const;                                                                // :67-69
};

struct Atoms {
    size_t n_atoms = 0;
    std::vector<int32_t> type_ids;   // 1-based
    std::vector<double> masses;      // per type
    std::vector<double> positions, velocities, forces;  // 3N, xyz interleaved (Matrix3xX memory image)
    SimulationBox sim_box{};
    double mass_i(size_t i) const { return masses[(size_t)type_ids[i] - 1]; }           // properties.rs:9-13
    size_t degress_of_freedom() const { return 3 * n_atoms; }                             // :41-43
    double temerature(double kinetic_energy) const;                                       // :28-30
    double pressure(double kinetic_energy, double virial_trace) const;                    // :61-65
    void start_velocities(double temperature, size_t seed);                               // velocities.rs:10-59
};

struct LennardJones {
    double epsilon, sigma, rcut;
    bool shift;
    double get_rcut() const { return rcut; }
};

// Device-backed drop-in for LJVOffsetManager (the new PotentialManagerKind variant).
class LJCudaManager {
   public:
    explicit LJCudaManager(double skin = 0.0, int device = 0) : skin_(skin), device_(device) {}
    ~LJCudaManager();
    LJCudaManager(const LJCudaManager &) = delete;
    LJCudaManager &operator=(const LJCudaManager &) = delete;
    // PairPotentialManager surface (potential.rs:148-193)
    bool is_empty() const { return table.empty(); }
    void insert(std::pair<int, int> key, const LennardJones &p) { table[key] = p; dirty_ = true; }
    const LennardJones *get(std::pair<int, int> key) const;
    double max_rcut() const;
    // PotentialManager trait, host-buffer form (potential.rs:13,15-33)
    double compute_potential(Atoms &atoms);
    double verlet_step_nve(Atoms &atoms, double dt);
    // device-resident form used by Simulation::run
    void attach(const Atoms &atoms);
    double compute(bool accumulate);
    void step_nve(double dt, int64_t nsteps, pisb_thermo *out);
    void step_nvt_nhc(double dt, int64_t nsteps, pisb_nhc &chain, int64_t first_step, int64_t total_steps, pisb_thermo *out,
                      double *nhc_energy);  // verlet_step_nvt_nhc x nsteps (potential.rs:35-58)
    void step_npt_mtk(double dt, int64_t nsteps, pisb_mtk &baro, pisb_nhc &chain, int64_t first_step, int64_t total_steps,
                      pisb_thermo *out, double *ext_energy, double *h9_trace);  // verlet_step_npt_mtk x nsteps (potential.rs:112-135)
    void download(Atoms &atoms, bool pos, bool vel, bool frc);
    // asynchronous position download of a dump step: begin takes a device-side snapshot and copies it on a second stream
    // while the next batch runs; the array is valid after end
    void download_begin(Atoms &atoms, bool pos, bool vel, bool frc);
    void download_end();
    void start_velocities(double temperature, size_t seed);  // Atoms::start_velocities (velocities.rs:10-15) on the device
    pisb_stats_t stats();
    std::map<std::pair<int, int>, LennardJones> table;

   private:
    void ensure_handle(const Atoms &atoms);
    void check(int rc);
    double skin_;
    int device_;
    pisb_t *h_ = nullptr;
    bool dirty_ = true;
    size_t n_types_ = 0;
};

// ---- readers ---------------------------------------------------------------------------------
struct StartVelocity {             // simulation_context.rs:14-32
    std::string group = "all";
    bool start_velocity = true;
    std::optional<double> start_temperature;
    std::optional<size_t> seed;
    std::optional<std::string> dist;
};
struct NHThermostatChainArgs { std::string name, group; double start_temperature, end_temperature, tau; };
struct MTKBarostatArgs { std::string name, group; double start_pressure, tau; };
struct PotentialArgs {
    std::vector<std::string> pair_style_args;
    size_t pair_style_line = 0;
    std::vector<std::vector<std::string>> pair_coeff_args;
    std::vector<size_t> pair_coeff_lines;
};
struct DumpArgs {                  // simulation_context.rs:78-102
    std::string name = "default_dump", group = "all", style = "atoms";
    size_t dump_step = 1;
    std::string file_name = "dump.lammpstrj";
};
struct SimulationContext {         // simulation_context.rs:105-131
    std::optional<Atoms> atoms;
    double timestep = 0.01;
    std::unique_ptr<LJCudaManager> mgr;
    std::optional<StartVelocity> starting_velocity;
    size_t steps = 100;
    std::optional<PotentialArgs> potential_args;
    std::optional<NHThermostatChainArgs> nh_chain_args;
    std::optional<MTKBarostatArgs> mtk_barostat_args;
    DumpArgs dump_args;
    double skin = 0.0;  // NEW knob (not a reference command): Verlet skin handed to the CUDA manager
    int device = 0;
    // `velocity ... create` is executed on the DEVICE once the atoms are resident (pisb_start_velocities): contextualize
    // records the request here instead of running the host generator (same numbers, see Atoms::start_velocities)
    bool device_velocities = false;
    bool velocities_pending = false;
};

// One script statement: looks the keyword up in the statement table and lets its reader consume the words.
// Returns false for an unknown keyword.
bool run_command(const std::string &command, const std::vector<std::string> &args, size_t line, SimulationContext &ctx);

class DumpTraj {                   // dump_traj.rs:12-75
   public:
    explicit DumpTraj(const DumpArgs &args);
    ~DumpTraj();
    void write_step(const Atoms &atoms, size_t step);

   private:
    FILE *out_ = nullptr;
    std::string path_;
};

enum class Ensemble { NVE, NVT, NPT };  // simulation.rs:118-133

struct Simulation {
    static void run(LJCudaManager &mgr, SimulationContext &ctx, FILE *thermo_out);  // simulation.rs:8-115 (NVE, NVT, NPT)
};

class System {                     // system.rs:35-183
   public:
    explicit System(std::string infile) : infile_(std::move(infile)) {}
    System &read();
    System &contextualize();
    void run(FILE *thermo_out = stdout);
    SimulationContext ctx;
    std::string describe() const;  // JSON summary of the parsed context (tests; `--check`)

   private:
    std::string infile_;
};

}  // namespace pis
