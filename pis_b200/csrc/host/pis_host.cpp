// pis_host.cpp -- see pis_host.hpp.  Cold path: text parsing, dump writing, the NVE run loop.
#include "pis_host.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace pis {

// ---- small helpers ------------------------------------------------------------------------------
static std::string trim(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}

static std::vector<std::string> split_whitespace(const std::string &s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}

std::string rust_display_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);  // shortest round-trip, fixed
    return std::string(buf, r.ptr);
}

// ArgsExt (src/extensions.rs:3-31)
static const std::string &get_required(const std::vector<std::string> &a, size_t i, size_t line) {
    if (i >= a.size()) throw PisError("MissingArgument", "Missing argument on line " + std::to_string(line));
    return a[i];
}

static int32_t parse_i32(const std::string &arg) {  // Rust str::parse::<i32>
    auto bad = [&](const char *why) {
        return PisError("IntParseError", "Error parsing integer number from string " + arg + ": " + why);
    };
    if (arg.empty()) throw bad("cannot parse integer from empty string");
    size_t p = 0;
    bool neg = false;
    if (arg[0] == '+' || arg[0] == '-') {
        neg = arg[0] == '-';
        p = 1;
        if (arg.size() == 1) throw bad("invalid digit found in string");
    }
    long long v = 0;
    for (; p < arg.size(); ++p) {
        if (arg[p] < '0' || arg[p] > '9') throw bad("invalid digit found in string");
        v = v * 10 + (arg[p] - '0');
        if (v > 2147483648LL) throw bad(neg ? "number too small to fit in target type" : "number too large to fit in target type");
    }
    if (neg) v = -v;
    if (v > 2147483647LL) throw bad("number too large to fit in target type");
    return (int32_t)v;
}

static double parse_f64(const std::string &arg) {  // Rust str::parse::<f64>
    auto bad = [&]() {
        return PisError("FloatParseError", "Error parsing floating number from string " + arg + ": invalid float literal");
    };
    if (arg.empty()) throw bad();
    std::string low;
    for (char c : arg) low.push_back((char)std::tolower((unsigned char)c));
    size_t p = (low[0] == '+' || low[0] == '-') ? 1 : 0;
    const std::string body = low.substr(p);
    if (body == "inf" || body == "infinity") return low[0] == '-' ? -INFINITY : INFINITY;
    if (body == "nan") return NAN;
    bool digit = false;
    for (char c : body) {
        if (std::isdigit((unsigned char)c)) digit = true;
        else if (c != '.' && c != 'e' && c != '+' && c != '-') throw bad();
    }
    if (!digit) throw bad();
    char *end = nullptr;
    const double v = std::strtod(arg.c_str(), &end);
    if (end != arg.c_str() + arg.size()) throw bad();
    return v;
}

static int32_t parse_int_at(const std::vector<std::string> &a, size_t i, size_t line) { return parse_i32(get_required(a, i, line)); }
static double parse_float_at(const std::vector<std::string> &a, size_t i, size_t line) { return parse_f64(get_required(a, i, line)); }

static size_t convert_to_usize(int32_t v, size_t line) {  // src/extensions.rs:33-44
    if (v < 0) throw PisError("NegativeValue", "Negative value " + std::to_string(v) + " not allowed on line: " + std::to_string(line));
    return (size_t)v;
}

// ---- SimulationBox -------------------------------------------------------------------------------
SimulationBox SimulationBox::make(const double m[9], const bool pbc_in[3]) {
    // nalgebra Matrix3::try_inverse (adjugate / determinant); element (r,c) at [c*3 + r]
#define M(r, c) m[((c)-1) * 3 + ((r)-1)]
    const double minor_m12_m23 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const double minor_m11_m23 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const double minor_m11_m22 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    const double det = M(1, 1) * minor_m12_m23 - M(1, 2) * minor_m11_m23 + M(1, 3) * minor_m11_m22;
    if (det == 0.0) throw PisError("SingularBox", "Box matrix should be invertible");
    SimulationBox b;
    std::memcpy(b.h, m, sizeof b.h);
#define O(r, c) b.h_inv[((c)-1) * 3 + ((r)-1)]
    O(1, 1) = minor_m12_m23 / det;
    O(1, 2) = (M(1, 3) * M(3, 2) - M(3, 3) * M(1, 2)) / det;
    O(1, 3) = (M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3)) / det;
    O(2, 1) = -minor_m11_m23 / det;
    O(2, 2) = (M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3)) / det;
    O(2, 3) = (M(1, 3) * M(2, 1) - M(2, 3) * M(1, 1)) / det;
    O(3, 1) = minor_m11_m22 / det;
    O(3, 2) = (M(1, 2) * M(3, 1) - M(3, 2) * M(1, 1)) / det;
    O(3, 3) = (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) / det;
#undef O
#undef M
    for (int d = 0; d < 3; ++d) b.pbc[d] = pbc_in[d];
    return b;
}

SimulationBox SimulationBox::from_lammps_data(double xlo, double xhi, double ylo, double yhi, double zlo, double zhi,
                                              double xy, double xz, double yz) {
    const double h[9] = {xhi - xlo, 0.0, 0.0, xy, yhi - ylo, 0.0, xz, yz, zhi - zlo};
    const bool pbc[3] = {true, true, true};
    return make(h, pbc);
}

double SimulationBox::volume() const {
#define M(r, c) h[((c)-1) * 3 + ((r)-1)]
    const double a = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const double b = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const double c = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    return std::fabs(M(1, 1) * a - M(1, 2) * b + M(1, 3) * c);
#undef M
}

// ---- Atoms ---------------------------------------------------------------------------------------
double Atoms::temerature(double kinetic_energy) const {
    return (2.0 * kinetic_energy) / ((double)degress_of_freedom() * KB_KJPERMOLEKELVIN);
}

double Atoms::pressure(double kinetic_energy, double virial_trace) const {
    return (2.0 * kinetic_energy + virial_trace) / (3.0 * sim_box.volume());
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// `velocity all create T seed` (velocities.rs:10-59).  The reference's stream (rand SmallRng +
// rand_distr Normal) is third-party and unpinned, so the Gaussian numbers come from the same id-keyed
// generator as pis_b200/decomposition.py (splitmix64 + Box-Muller); drift removal and rescaling follow
// the reference.
void Atoms::start_velocities(double temperature, size_t seed) {
    const uint64_t base = splitmix64((uint64_t)seed * 0xD1342543DE82EF95ULL + 12345ULL);
    for (size_t i = 0; i < n_atoms; ++i) {
        const double sigma = std::sqrt(KB_KJPERMOLEKELVIN * temperature / mass_i(i));
        if (!(sigma >= 0.0) || !std::isfinite(sigma))
            throw PisError("InvalidDistribution", "Invalid normal distribution parameters: standard deviation is invalid");
        for (int c = 0; c < 3; ++c) {
            const uint64_t k = splitmix64((uint64_t)i * 6 + 2 * c + base), k2 = splitmix64((uint64_t)i * 6 + 2 * c + 1 + base);
            const double u1 = ((double)(k >> 11) + 1.0) * (1.0 / 9007199254740993.0);
            const double u2 = (double)(k2 >> 11) * (1.0 / 9007199254740992.0);
            velocities[3 * i + c] = std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2) * sigma;
        }
    }
    double total_mass = 0.0, mom[3] = {0.0, 0.0, 0.0};
    for (size_t i = 0; i < n_atoms; ++i) {
        const double m = mass_i(i);
        total_mass += m;
        for (int c = 0; c < 3; ++c) mom[c] += velocities[3 * i + c] * m;
    }
    for (size_t i = 0; i < n_atoms; ++i)
        for (int c = 0; c < 3; ++c) velocities[3 * i + c] -= (mom[c] / total_mass) * 1.0;
    double ek = 0.0;
    for (size_t i = 0; i < n_atoms; ++i) {
        const double *v = &velocities[3 * i];
        ek += 0.5 * mass_i(i) * ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    }
    const double lambda = std::sqrt(temperature / temerature(ek));
    for (double &v : velocities) v *= lambda;
}

// ---- LJCudaManager -------------------------------------------------------------------------------
LJCudaManager::~LJCudaManager() {
    if (h_) pisb_destroy(h_);
}

void LJCudaManager::check(int rc) {
    if (rc != PISB_OK) throw PisError("Device", std::string("pisb error ") + std::to_string(rc) + ": " + pisb_last_error(h_));
}

const LennardJones *LJCudaManager::get(std::pair<int, int> key) const {
    auto it = table.find(key);
    return it == table.end() ? nullptr : &it->second;
}

double LJCudaManager::max_rcut() const {  // potential.rs:170-179
    double m = 0.0;
    for (auto &kv : table)
        if (m < kv.second.rcut) m = kv.second.rcut;
    return m;
}

void LJCudaManager::ensure_handle(const Atoms &atoms) {
    size_t nt = atoms.masses.size();
    for (auto &kv : table) nt = std::max(nt, (size_t)std::max(kv.first.first, kv.first.second));
    if (nt < 1) nt = 1;
    if (h_ && !dirty_ && nt == n_types_) return;
    if (h_) pisb_destroy(h_);
    h_ = nullptr;
    std::vector<double> mass(nt, 0.0), eps(nt * nt, 0.0), sig(nt * nt, 0.0), rc(nt * nt, 0.0);
    std::vector<unsigned char> present(nt * nt, 0);
    for (size_t t = 0; t < atoms.masses.size(); ++t) mass[t] = atoms.masses[t];
    bool shift = true;
    for (auto &kv : table) {
        if (kv.first.first < 1 || kv.first.second < 1) continue;
        const size_t k = (size_t)(kv.first.first - 1) * nt + (size_t)(kv.first.second - 1);
        eps[k] = kv.second.epsilon;
        sig[k] = kv.second.sigma;
        rc[k] = kv.second.rcut;
        present[k] = 1;
        shift = kv.second.shift;
    }
    int rcode = pisb_create(device_, (int)nt, mass.data(), eps.data(), sig.data(), rc.data(), present.data(), shift ? 1 : 0,
                            skin_, &h_);
    if (rcode != PISB_OK) throw PisError("Device", std::string("pisb error ") + std::to_string(rcode) + ": " + pisb_last_error(nullptr));
    dirty_ = false;
    n_types_ = nt;
}

void LJCudaManager::attach(const Atoms &atoms) {
    ensure_handle(atoms);
    int pbc[3] = {atoms.sim_box.pbc[0], atoms.sim_box.pbc[1], atoms.sim_box.pbc[2]};
    check(pisb_set_box(h_, atoms.sim_box.h, atoms.sim_box.h_inv, pbc));
    check(pisb_upload(h_, (int64_t)atoms.n_atoms, atoms.positions.data(), atoms.velocities.data(), atoms.forces.data(),
                      atoms.type_ids.data()));
}

double LJCudaManager::compute(bool accumulate) {
    double pe = 0.0;
    check(pisb_compute(h_, accumulate ? 1 : 0, &pe));
    return pe;
}

void LJCudaManager::step_nve(double dt, int64_t nsteps, pisb_thermo *out) { check(pisb_step_nve(h_, dt, nsteps, out)); }

void LJCudaManager::step_nvt_nhc(double dt, int64_t nsteps, pisb_nhc &chain, int64_t first_step, int64_t total_steps,
                                 pisb_thermo *out, double *nhc_energy) {
    check(pisb_step_nvt_nhc(h_, dt, nsteps, &chain, first_step, total_steps, out, nhc_energy));
}

void LJCudaManager::step_npt_mtk(double dt, int64_t nsteps, pisb_mtk &baro, pisb_nhc &chain, int64_t first_step, int64_t total_steps,
                                 pisb_thermo *out, double *ext_energy, double *h9_trace) {
    check(pisb_step_npt_mtk(h_, dt, nsteps, &baro, &chain, first_step, total_steps, out, ext_energy, h9_trace));
}

void LJCudaManager::download(Atoms &atoms, bool pos, bool vel, bool frc) {
    check(pisb_download(h_, pos ? atoms.positions.data() : nullptr, vel ? atoms.velocities.data() : nullptr,
                        frc ? atoms.forces.data() : nullptr));
}

void LJCudaManager::download_begin(Atoms &atoms, bool pos, bool vel, bool frc) {
    check(pisb_download_begin(h_, pos ? atoms.positions.data() : nullptr, vel ? atoms.velocities.data() : nullptr,
                              frc ? atoms.forces.data() : nullptr));
}

void LJCudaManager::download_end() { check(pisb_download_end(h_)); }

pisb_stats_t LJCudaManager::stats() {
    pisb_stats_t s{};
    check(pisb_stats(h_, &s));
    return s;
}

double LJCudaManager::compute_potential(Atoms &atoms) {  // adds into atoms.forces, returns PE
    attach(atoms);
    const double pe = compute(true);
    download(atoms, false, false, true);
    return pe;
}

double LJCudaManager::verlet_step_nve(Atoms &atoms, double dt) {
    ensure_handle(atoms);
    int pbc[3] = {atoms.sim_box.pbc[0], atoms.sim_box.pbc[1], atoms.sim_box.pbc[2]};
    check(pisb_set_box(h_, atoms.sim_box.h, atoms.sim_box.h_inv, pbc));
    double pe = 0.0;
    check(pisb_verlet_step_nve_host(h_, (int64_t)atoms.n_atoms, atoms.positions.data(), atoms.velocities.data(),
                                    atoms.forces.data(), atoms.type_ids.data(), dt, &pe));
    return pe;
}

// ---- commands (src/readers/input_file/commands.rs) ---------------------------------------------------
static void run_timestep(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) { ctx.timestep = parse_float_at(a, 0, line); }

static void run_runsteps(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {
    ctx.steps = convert_to_usize(parse_int_at(a, 0, line), line);
}

static void run_velocity(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {  // :78-144
    size_t read_args = 0;
    StartVelocity sv;
    sv.group = get_required(a, read_args, line);
    ++read_args;
    const std::string style = get_required(a, read_args, line);
    ++read_args;
    if (style == "create") {
        sv.start_temperature = parse_float_at(a, read_args, line);
        ++read_args;
        int32_t seed = 0;
        bool ok = true;
        try {
            seed = parse_int_at(a, read_args, line);
        } catch (const PisError &) {
            ok = false;
        }
        if (ok) {
            ++read_args;
            sv.seed = convert_to_usize(seed, line);
        } else {
            sv.seed = 0;
        }
    } else {
        throw PisError("InvalidArgument", "Invalid argument: " + style + " at line: " + std::to_string(line));
    }
    while (read_args < a.size()) {
        const std::string keyword = a[read_args++];
        if (keyword == "dist") {
            const std::string arg = get_required(a, read_args, line);
            ++read_args;
            if (arg == "uniform" || arg == "gaussian") sv.dist = arg;
            else throw PisError("InvalidArgument", "Invalid argument: " + arg + " at line: " + std::to_string(line));
        } else {
            throw PisError("InvalidArgument", "Invalid argument: " + keyword + " at line: " + std::to_string(line));
        }
    }
    ctx.starting_velocity = sv;
}

static void run_read_data(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {  // :147-384
    const std::string path = get_required(a, 0, line);
    std::ifstream file(path);
    if (!file) throw PisError("InputFileError", "Failed to open input file '" + path + "': No such file or directory (os error 2))");
    std::string section;
    size_t n_atoms = 0;
    double xlo = 0.0, xhi = 1.0, ylo = 0.0, yhi = 1.0, zlo = 0.0, zhi = 1.0;
    std::vector<double> masses, positions, velocities;
    std::vector<int32_t> type_ids;
    bool start_velocities = true;
    auto mgr = std::make_unique<LJCudaManager>(ctx.skin, ctx.device);
    std::string raw;
    size_t line_num = 0;
    for (; std::getline(file, raw); ++line_num) {
        const std::string l = trim(raw);
        if (l.empty() || l[0] == '#') continue;
        const std::vector<std::string> ls = split_whitespace(l);
        const std::string &first = get_required(ls, 0, line_num);
        if (first == "Masses" || first == "Atoms" || first == "PairCoeffs") {
            section = first;
            continue;
        }
        if (first == "Velocities") {
            start_velocities = false;
            section = first;
            continue;
        }
        if (ls.size() > 1) {
            if (ls[1] == "atoms") {
                n_atoms = convert_to_usize(parse_int_at(ls, 0, line_num), line_num);
                type_ids.assign(n_atoms, 0);
                positions.assign(3 * n_atoms, 0.0);
                velocities.assign(3 * n_atoms, 0.0);
                continue;
            }
            if (ls[1] == "atom") {
                masses.resize(convert_to_usize(parse_int_at(ls, 0, line_num), line_num), 0.0);
                continue;
            }
        }
        if (ls.size() > 2) {
            if (ls[2] == "xlo") { xlo = parse_float_at(ls, 0, line_num); xhi = parse_float_at(ls, 1, line_num); continue; }
            if (ls[2] == "ylo") { ylo = parse_float_at(ls, 0, line_num); yhi = parse_float_at(ls, 1, line_num); continue; }
            if (ls[2] == "zlo") { zlo = parse_float_at(ls, 0, line_num); zhi = parse_float_at(ls, 1, line_num); continue; }
        }
        if (section == "Masses") {
            const size_t type_id = convert_to_usize(parse_int_at(ls, 0, line_num), line_num);
            const double mass = parse_float_at(ls, 1, line_num);
            if (type_id < 1) throw PisError("InvalidAtomType", "Atom type " + std::to_string(type_id) + " out of range");
            if (type_id > masses.size()) throw PisError("InvalidAtomType", "Atom type " + std::to_string(type_id) + " out of range");  // reference: panic (OOB)
            masses[type_id - 1] = mass;
        } else if (section == "PairCoeffs") {
            // "1 0.238 3.405 8.5" -> (1,1); the "i j eps sigma rc" form is unreachable for integer j because
            // token 1 parses as a float first (reference quirk, commands.rs:271-291) -- reproduced.
            const int i = (int)convert_to_usize(parse_int_at(ls, 0, line_num), line_num);
            bool eps_ok = true;
            double epsilon = 0.0;
            try {
                epsilon = parse_float_at(ls, 1, line_num);
            } catch (const PisError &) {
                eps_ok = false;
            }
            if (eps_ok) {
                const double sigma = parse_float_at(ls, 2, line_num);
                double rcut;
                try { rcut = parse_float_at(ls, 3, line_num); } catch (const PisError &) { rcut = 2.5 * sigma; }
                mgr->insert({i, i}, LennardJones{epsilon, sigma, rcut, true});
            } else {
                const int j = (int)convert_to_usize(parse_int_at(ls, 1, line_num), line_num);
                epsilon = parse_float_at(ls, 2, line_num);
                const double sigma = parse_float_at(ls, 3, line_num);
                double rcut;
                try { rcut = parse_float_at(ls, 4, line_num); } catch (const PisError &) { rcut = 2.5 * sigma; }
                mgr->insert({i, j}, LennardJones{epsilon, sigma, rcut, true});
            }
            continue;
        } else if (section == "Atoms") {
            size_t id = convert_to_usize(parse_int_at(ls, 0, line_num), line_num);
            if (id == 0 || id > n_atoms)
                throw PisError("AtomCountMismatch", "Atom count mismatch: expected " + std::to_string(n_atoms) + ", found " + std::to_string(id));
            --id;
            type_ids[id] = (int32_t)convert_to_usize(parse_int_at(ls, 1, line_num), line_num);
            positions[3 * id] = parse_float_at(ls, 2, line_num);
            positions[3 * id + 1] = parse_float_at(ls, 3, line_num);
            positions[3 * id + 2] = parse_float_at(ls, 4, line_num);
        } else if (section == "Velocities") {
            size_t id = convert_to_usize(parse_int_at(ls, 0, line_num), line_num);
            if (id == 0 || id > n_atoms)
                throw PisError("AtomCountMismatch", "Atom count mismatch: expected " + std::to_string(n_atoms) + ", found " + std::to_string(id));
            --id;
            velocities[3 * id] = parse_float_at(ls, 1, line_num);
            velocities[3 * id + 1] = parse_float_at(ls, 2, line_num);
            velocities[3 * id + 2] = parse_float_at(ls, 3, line_num);
        }
    }
    Atoms atoms;
    atoms.n_atoms = n_atoms;
    atoms.type_ids = std::move(type_ids);
    atoms.masses = std::move(masses);
    atoms.positions = std::move(positions);
    atoms.velocities = std::move(velocities);
    atoms.forces.assign(3 * n_atoms, 0.0);
    atoms.sim_box = SimulationBox::from_lammps_data(xlo, xhi, ylo, yhi, zlo, zhi, 0.0, 0.0, 0.0);
    ctx.atoms = std::move(atoms);
    if (!mgr->is_empty()) ctx.mgr = std::move(mgr);
    if (ctx.starting_velocity) {
        ctx.starting_velocity->start_velocity = start_velocities;
    } else {
        StartVelocity sv;
        sv.start_velocity = start_velocities;
        ctx.starting_velocity = sv;
    }
}

static void run_fix(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {  // :387-459
    size_t read_args = 0;
    const std::string name = get_required(a, read_args++, line);
    const std::string group = get_required(a, read_args++, line);
    const std::string style = get_required(a, read_args++, line);
    while (read_args < a.size()) {
        const std::string keyword = a[read_args++];
        if (keyword == "temp") {
            const double t0 = parse_float_at(a, read_args++, line);
            const double t1 = parse_float_at(a, read_args++, line);
            const double tau = parse_float_at(a, read_args++, line);
            if (style == "npt" || style == "nvt") ctx.nh_chain_args = NHThermostatChainArgs{name, group, t0, t1, tau};
        } else if (keyword == "iso") {
            const double p0 = parse_float_at(a, read_args++, line);
            (void)parse_float_at(a, read_args++, line);
            const double tau = parse_float_at(a, read_args++, line);
            if (style == "npt") ctx.mtk_barostat_args = MTKBarostatArgs{name, group, p0, tau};
        } else {
            throw PisError("InvalidArgument", "Invalid argument: " + keyword + " at line: " + std::to_string(line));
        }
    }
}

static void run_pair_style(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {
    PotentialArgs p;
    p.pair_style_line = line;
    p.pair_style_args = a;
    ctx.potential_args = p;
}

static void run_pair_coeff(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {
    if (!ctx.potential_args)
        throw PisError("PotentialNotInitialized", "Potential manager not initialized - missing pair_style or pair_coeff commands");
    ctx.potential_args->pair_coeff_args.push_back(a);
    ctx.potential_args->pair_coeff_lines.push_back(line);
}

static void run_dump(const std::vector<std::string> &a, size_t line, SimulationContext &ctx) {  // :484-496
    size_t r = 0;
    ctx.dump_args.name = get_required(a, r++, line);
    ctx.dump_args.group = get_required(a, r++, line);
    ctx.dump_args.style = get_required(a, r++, line);
    ctx.dump_args.dump_step = convert_to_usize(parse_int_at(a, r++, line), line);
    ctx.dump_args.file_name = get_required(a, r, line);
}

bool run_command(const std::string &command, const std::vector<std::string> &args, size_t line, SimulationContext &ctx) {
    if (command == "timestep") run_timestep(args, line, ctx);
    else if (command == "run") run_runsteps(args, line, ctx);
    else if (command == "velocity") run_velocity(args, line, ctx);
    else if (command == "read_data") run_read_data(args, line, ctx);
    else if (command == "fix") run_fix(args, line, ctx);
    else if (command == "pair_style") run_pair_style(args, line, ctx);
    else if (command == "pair_coeff") run_pair_coeff(args, line, ctx);
    else if (command == "dump") run_dump(args, line, ctx);
    else return false;
    return true;
}

// ---- DumpTraj (src/writers/dump_traj.rs) ---------------------------------------------------------------
DumpTraj::DumpTraj(const DumpArgs &args) : path_(args.file_name) {
    out_ = std::fopen(path_.c_str(), "w");
    if (!out_) throw PisError("DumpCreateError", "Failed to create trajectory file '" + path_ + "': " + std::strerror(errno));
    static char *buf = nullptr;
    (void)buf;
    std::setvbuf(out_, nullptr, _IOFBF, 1 << 20);
}

DumpTraj::~DumpTraj() {
    if (out_) std::fclose(out_);
}

void DumpTraj::write_step(const Atoms &atoms, size_t step) {
    std::string s;
    s.reserve(64 * atoms.n_atoms + 256);
    s += "ITEM: TIMESTEP\n" + std::to_string(step) + "\n";
    s += "ITEM: NUMBER OF ATOMS\n" + std::to_string(atoms.n_atoms) + "\n";
    s += "ITEM: BOX BOUNDS pp pp pp\n";
    for (int d = 0; d < 3; ++d) s += "0 " + rust_display_f64(atoms.sim_box.h[4 * d]) + "\n";
    s += "ITEM: ATOMS id type x y z\n";
    for (size_t i = 0; i < atoms.n_atoms; ++i) {
        s += std::to_string(i + 1);
        s += ' ';
        s += std::to_string(atoms.type_ids[i]);
        for (int c = 0; c < 3; ++c) {
            s += ' ';
            s += rust_display_f64(atoms.positions[3 * i + c]);
        }
        s += '\n';
    }
    if (std::fwrite(s.data(), 1, s.size(), out_) != s.size())
        throw PisError("DumpWriteError", "Failed to write trajectory file '" + path_ + "': " + std::strerror(errno));
}

// ---- Simulation::run (src/simulation.rs:8-115): NVE, NVT and NPT arms ---------------------------------------
void Simulation::run(LJCudaManager &mgr, SimulationContext &ctx, FILE *out) {
    // Ensemble::from_ctx (simulation.rs:125-132): (thermostat, barostat) => NPT, (thermostat, -) => NVT, else NVE
    const bool nvt = ctx.nh_chain_args.has_value();
    const bool npt = nvt && ctx.mtk_barostat_args.has_value();
    pisb_nhc chain{};
    if (nvt) pisb_nhc_init(&chain, ctx.nh_chain_args->start_temperature, ctx.nh_chain_args->end_temperature, ctx.nh_chain_args->tau);
    if (!ctx.atoms) throw PisError("NoAtomsDefined", "No atoms defined in input file");
    Atoms &atoms = *ctx.atoms;
    pisb_mtk baro{};
    if (npt) {
        // MTKBarostat::new_from_args (npt.rs:67-88): target = start_pressure * identity (commands.rs:439), T = thermostat start
        const double p = ctx.mtk_barostat_args->start_pressure;
        const double target[9] = {p, 0, 0, 0, p, 0, 0, 0, p};
        pisb_mtk_init(&baro, target, ctx.mtk_barostat_args->tau, (int64_t)atoms.n_atoms, ctx.nh_chain_args->start_temperature);
    }
    const double dt = ctx.timestep;
    const size_t steps = ctx.steps, dump_step = ctx.dump_args.dump_step;
    DumpTraj dumper(ctx.dump_args);
    dumper.write_step(atoms, 0);
    const double first_potential = mgr.compute_potential(atoms);  // uploads; forces stay resident too
    std::fprintf(out, "0 %s\n", rust_display_f64(first_potential).c_str());
    std::vector<pisb_thermo> th;
    std::vector<double> ext_e, h_trace;
    // Dump steps (NVE / NVT, where the box does not change): the positions are snapshotted on the device and copied back on a
    // second stream while the NEXT batch of steps runs; the frame is written when that batch returns.  The position array
    // is page-locked for the run so the copy is a true asynchronous DMA.
    const bool async_dump = !npt && dump_step > 0;
    const bool pos_pinned = async_dump && pisb_host_register(atoms.positions.data(), atoms.positions.size() * sizeof(double)) == PISB_OK;
    size_t pending_dump = 0;  // step whose positions are in flight (0 = none)
    auto flush_dump = [&]() {
        if (pending_dump == 0) return;
        mgr.download_end();
        dumper.write_step(atoms, pending_dump);
        pending_dump = 0;
    };
    size_t i = 0;
    while (i < steps) {
        // run up to the next dump step on the device; only thermo scalars come back per step
        size_t chunk = steps - i;
        if (dump_step > 0) chunk = std::min(chunk, dump_step - (i % dump_step));
        chunk = std::min<size_t>(chunk, 1000);
        th.resize(chunk);
        ext_e.assign(chunk, 0.0);
        if (npt) {
            h_trace.assign(9 * chunk, 0.0);
            mgr.step_npt_mtk(dt, (int64_t)chunk, baro, chain, (int64_t)i, (int64_t)steps, th.data(), ext_e.data(), h_trace.data());
        } else if (nvt) {
            mgr.step_nvt_nhc(dt, (int64_t)chunk, chain, (int64_t)i, (int64_t)steps, th.data(), ext_e.data());
        } else {
            mgr.step_nve(dt, (int64_t)chunk, th.data());
        }
        flush_dump();  // the frame of the previous dump step travelled while this batch ran
        for (size_t k = 0; k < chunk; ++k) {
            const size_t step = i + k + 1;
            if (npt) {  // scale_box changed atoms.sim_box (transformations.rs:8-13): pressure and dump bounds use the step's box
                bool pbc[3] = {atoms.sim_box.pbc[0], atoms.sim_box.pbc[1], atoms.sim_box.pbc[2]};
                atoms.sim_box = SimulationBox::make(&h_trace[9 * k], pbc);
            }
            if (dump_step > 0 && step % dump_step == 0) {  // batches end on dump steps, so the device state IS this step's
                if (async_dump) {
                    pending_dump = step;
                    mgr.download_begin(atoms, true, false, false);
                } else {
                    mgr.download(atoms, true, false, false);
                    dumper.write_step(atoms, step);
                }
            }
            const double ke = th[k].ke, pe = th[k].pe;
            // compute_hamiltonian (simulation.rs:90-115): NVT adds the thermostat's kinetic + potential energy, NPT also
            // the barostat's
            std::fprintf(out, "%zu %.3f %.3f %.3f %.3f %.3f\n", step, pe, ke, pe + ke + ext_e[k], atoms.temerature(ke),
                         atoms.pressure(ke, th[k].virial_ref));
        }
        i += chunk;
    }
    flush_dump();
    if (pos_pinned) pisb_host_unregister(atoms.positions.data());
    mgr.download(atoms, true, true, true);
    std::fflush(out);
}

// ---- System (src/system.rs) ----------------------------------------------------------------------------
System &System::read() {
    std::ifstream file(infile_);
    if (!file) throw PisError("InputFileError", "Failed to open input file '" + infile_ + "': No such file or directory (os error 2))");
    std::string raw;
    size_t line_num = 0;
    while (std::getline(file, raw)) {
        ++line_num;
        const std::string l = trim(raw);
        if (l.empty() || l[0] == '#') continue;
        const size_t hash = l.find('#');
        const std::string uncommented = trim(hash == std::string::npos ? l : l.substr(0, hash));
        const std::vector<std::string> ls = split_whitespace(uncommented);
        if (ls.size() < 2) continue;
        const std::vector<std::string> args(ls.begin() + 1, ls.end());
        if (!run_command(ls[0], args, line_num, ctx))
            throw PisError("UnknownCommand", "Invalid command " + ls[0] + " found line: " + std::to_string(line_num));
    }
    return *this;
}

System &System::contextualize() {
    if (ctx.atoms) {
        if (ctx.atoms->n_atoms == 0) throw PisError("NoAtomsDefined", "No atoms defined in input file");
        if (ctx.starting_velocity && ctx.starting_velocity->start_velocity) {
            const double t = ctx.starting_velocity->start_temperature.value_or(300.0);
            const size_t seed = ctx.starting_velocity->seed.value_or(0);
            ctx.atoms->start_velocities(t, seed);
        }
    }
    if (ctx.potential_args) {
        const PotentialArgs &pa = *ctx.potential_args;
        const std::string &style = get_required(pa.pair_style_args, 0, pa.pair_style_line);
        if (style == "lj/cut") {
            const double global_cutoff = parse_float_at(pa.pair_style_args, 1, pa.pair_style_line);
            auto mgr = std::make_unique<LJCudaManager>(ctx.skin, ctx.device);
            for (size_t k = 0; k < pa.pair_coeff_args.size(); ++k) {
                const auto &pc = pa.pair_coeff_args[k];
                const size_t cl = pa.pair_coeff_lines[k];
                const int i = (int)convert_to_usize(parse_int_at(pc, 0, cl), cl);
                const int j = (int)convert_to_usize(parse_int_at(pc, 1, cl), cl);
                const double epsilon = parse_float_at(pc, 2, cl);
                const double sigma = parse_float_at(pc, 3, cl);
                double local_rcut;
                try { local_rcut = parse_float_at(pc, 4, cl); } catch (const PisError &) { local_rcut = global_cutoff; }
                mgr->insert({i, j}, LennardJones{epsilon, sigma, local_rcut, true});  // key stored as given (:162)
            }
            ctx.mgr = std::move(mgr);
        } else {
            throw PisError("UnknownPairStyle", "Unknown pair style: '" + style + "'");
        }
    }
    return *this;
}

void System::run(FILE *thermo_out) {
    if (!ctx.mgr) throw PisError("PotentialNotInitialized", "Potential manager not initialized - missing pair_style or pair_coeff commands");
    Simulation::run(*ctx.mgr, ctx, thermo_out);
}

std::string System::describe() const {
    std::ostringstream o;
    o << "{\"timestep\": " << rust_display_f64(ctx.timestep) << ", \"steps\": " << ctx.steps;
    o << ", \"dump\": {\"name\": \"" << ctx.dump_args.name << "\", \"group\": \"" << ctx.dump_args.group << "\", \"style\": \""
      << ctx.dump_args.style << "\", \"dump_step\": " << ctx.dump_args.dump_step << ", \"file_name\": \"" << ctx.dump_args.file_name << "\"}";
    o << ", \"ensemble\": \"" << (ctx.nh_chain_args ? (ctx.mtk_barostat_args ? "NPT" : "NVT") : "NVE") << "\"";
    if (ctx.starting_velocity) {
        const auto &sv = *ctx.starting_velocity;
        o << ", \"velocity\": {\"group\": \"" << sv.group << "\", \"start_velocity\": " << (sv.start_velocity ? "true" : "false");
        if (sv.start_temperature) o << ", \"temperature\": " << rust_display_f64(*sv.start_temperature);
        if (sv.seed) o << ", \"seed\": " << *sv.seed;
        if (sv.dist) o << ", \"dist\": \"" << *sv.dist << "\"";
        o << "}";
    }
    if (ctx.atoms) {
        const Atoms &a = *ctx.atoms;
        o << ", \"atoms\": {\"n_atoms\": " << a.n_atoms << ", \"n_types\": " << a.masses.size() << ", \"box\": ["
          << rust_display_f64(a.sim_box.h[0]) << ", " << rust_display_f64(a.sim_box.h[4]) << ", " << rust_display_f64(a.sim_box.h[8]) << "]";
        o << ", \"first_positions\": [";
        for (size_t i = 0; i < std::min<size_t>(a.n_atoms, 4) * 3; ++i) o << (i ? ", " : "") << rust_display_f64(a.positions[i]);
        o << "], \"first_velocities\": [";
        for (size_t i = 0; i < std::min<size_t>(a.n_atoms, 4) * 3; ++i) o << (i ? ", " : "") << rust_display_f64(a.velocities[i]);
        o << "], \"masses\": [";
        for (size_t i = 0; i < a.masses.size(); ++i) o << (i ? ", " : "") << rust_display_f64(a.masses[i]);
        o << "], \"types_head\": [";
        for (size_t i = 0; i < std::min<size_t>(a.n_atoms, 8); ++i) o << (i ? ", " : "") << a.type_ids[i];
        o << "]}";
    }
    o << ", \"potential\": ";
    if (ctx.mgr) {
        o << "{\"max_rcut\": " << rust_display_f64(ctx.mgr->max_rcut()) << ", \"pairs\": [";
        bool firstp = true;
        for (auto &kv : ctx.mgr->table) {
            o << (firstp ? "" : ", ") << "{\"i\": " << kv.first.first << ", \"j\": " << kv.first.second << ", \"epsilon\": "
              << rust_display_f64(kv.second.epsilon) << ", \"sigma\": " << rust_display_f64(kv.second.sigma) << ", \"rcut\": "
              << rust_display_f64(kv.second.rcut) << "}";
            firstp = false;
        }
        o << "]}";
    } else {
        o << "null";
    }
    o << "}";
    return o.str();
}

}  // namespace pis
