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

void LJCudaManager::start_velocities(double temperature, size_t seed) {
    const int rc = pisb_start_velocities(h_, temperature, (uint64_t)seed);
    if (rc == PISB_ERR_INVALID && h_)  // the reference's error for a bad variance (velocities.rs:25-26, errors.rs)
        throw PisError("InvalidDistribution", pisb_last_error(h_));
    check(rc);
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
    // The reference prints "During force calculation between i and j atoms, potential was missing" for every candidate pair of
    // every step (lennard_jones.rs:216-222) and skips the pair; here the pairs are skipped the same way and the host says so once.
    pisb_stats_t st{};
    if (pisb_stats(h_, &st) == PISB_OK && st.missing_type_pairs > 0)
        std::fprintf(stderr, "warning: %lld type pair(s) with atoms present have no pair potential: their interactions are skipped\n",
                     (long long)st.missing_type_pairs);
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

// =====================================================================================================
// Input script and LAMMPS-data readers.
//
// Written from the FILE FORMATS the reference accepts (its example/ inputs, the cases of src/tests/command_tests.rs and
// the quirks listed in SURVEY appendix A), not from its parser: a typed token cursor, a statement table for the script
// and a line classifier + per-section row readers for the data file.  What must match the reference is behaviour --
// which inputs are accepted, what they mean, which error (kind and text) a bad input raises -- and tests/test_cli.py
// pins that through `--check`.
// =====================================================================================================

// One line's words, consumed left to right with typed accessors.  Every accessor that needs a word raises
// MissingArgument for the line when none is left; the maybe_* forms consume a word only if it has the wanted type.
class Words {
   public:
    Words(std::vector<std::string> words, size_t line) : w_(std::move(words)), line_(line) {}
    bool done() const { return at_ >= w_.size(); }
    size_t left() const { return w_.size() - at_; }
    size_t line() const { return line_; }
    const std::string &word() { return get_required(w_, at_++, line_); }
    double real() { return parse_f64(word()); }
    size_t count() { return convert_to_usize(parse_i32(word()), line_); }
    // the word `ahead` positions further on, or nullptr
    const std::string *peek(size_t ahead = 0) const { return at_ + ahead < w_.size() ? &w_[at_ + ahead] : nullptr; }
    std::optional<double> maybe_real() {
        if (done()) return std::nullopt;
        try {
            const double v = parse_f64(w_[at_]);
            ++at_;
            return v;
        } catch (const PisError &) {
            return std::nullopt;
        }
    }
    std::optional<size_t> maybe_count() {
        if (done()) return std::nullopt;
        int32_t v;
        try {
            v = parse_i32(w_[at_]);
        } catch (const PisError &) {
            return std::nullopt;
        }
        ++at_;
        return convert_to_usize(v, line_);  // an integer that is there but negative is an error, not "absent"
    }
    std::vector<std::string> rest() {
        std::vector<std::string> out(w_.begin() + (std::ptrdiff_t)std::min(at_, w_.size()), w_.end());
        at_ = w_.size();
        return out;
    }
    [[noreturn]] void reject(const std::string &what) const {
        throw PisError("InvalidArgument", "Invalid argument: " + what + " at line: " + std::to_string(line_));
    }

   private:
    std::vector<std::string> w_;
    size_t at_ = 0, line_;
};

// A trailing `keyword value value ...` clause of a statement: the keyword and how many reals follow it.
struct Clause {
    const char *keyword;
    int reals;
};

// Reads `keyword v1 .. vn` clauses until the line ends; unknown keywords are rejected.  store(k, values) receives the
// index of the clause in the table.
template <typename Store>
static void read_clauses(Words &w, const Clause *table, size_t n_table, Store store) {
    while (!w.done()) {
        const std::string key = w.word();
        size_t k = 0;
        while (k < n_table && key != table[k].keyword) ++k;
        if (k == n_table) w.reject(key);
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r < table[k].reals; ++r) v[r] = w.real();
        store(k, v);
    }
}

// ---- script statements ------------------------------------------------------------------------------
// timestep <dt>
static void stmt_timestep(Words &w, SimulationContext &ctx) { ctx.timestep = w.real(); }

// run <steps>
static void stmt_run(Words &w, SimulationContext &ctx) { ctx.steps = w.count(); }

// velocity <group> create <T> [<seed>] [dist uniform|gaussian]
//   the seed is optional: a fourth word that is not an integer leaves the seed at 0 and is read as a keyword
static void stmt_velocity(Words &w, SimulationContext &ctx) {
    StartVelocity sv;
    sv.group = w.word();
    const std::string style = w.word();
    if (style != "create") w.reject(style);
    sv.start_temperature = w.real();
    sv.seed = w.maybe_count().value_or(0);
    while (!w.done()) {
        const std::string key = w.word();
        if (key != "dist") w.reject(key);
        const std::string kind = w.word();
        if (kind != "uniform" && kind != "gaussian") w.reject(kind);
        sv.dist = kind;
    }
    ctx.starting_velocity = sv;  // replaces whatever a data file recorded before (start_velocity back to true)
}

// fix <name> <group> <style> [temp <T0> <T1> <tau>] [iso <P0> <P1> <tau>]
//   temp counts for the nvt and npt styles, iso for npt; any other style parses and changes nothing
static void stmt_fix(Words &w, SimulationContext &ctx) {
    const std::string name = w.word(), group = w.word(), style = w.word();
    static const Clause clauses[] = {{"temp", 3}, {"iso", 3}};
    const bool thermostat = style == "nvt" || style == "npt", barostat = style == "npt";
    read_clauses(w, clauses, 2, [&](size_t k, const double *v) {
        if (k == 0 && thermostat) ctx.nh_chain_args = NHThermostatChainArgs{name, group, v[0], v[1], v[2]};
        if (k == 1 && barostat) ctx.mtk_barostat_args = MTKBarostatArgs{name, group, v[0], v[2]};
    });
}

// pair_style <style> <args...>      kept verbatim; interpreted once the whole script has been read
static void stmt_pair_style(Words &w, SimulationContext &ctx) {
    PotentialArgs p;
    p.pair_style_line = w.line();
    p.pair_style_args = w.rest();
    ctx.potential_args = p;
}

// pair_coeff <i> <j> <eps> <sigma> [<rc>]      needs a pair_style before it
static void stmt_pair_coeff(Words &w, SimulationContext &ctx) {
    if (!ctx.potential_args)
        throw PisError("PotentialNotInitialized", "Potential manager not initialized - missing pair_style or pair_coeff commands");
    ctx.potential_args->pair_coeff_lines.push_back(w.line());
    ctx.potential_args->pair_coeff_args.push_back(w.rest());
}

// dump <name> <group> <style> <every> <file>
static void stmt_dump(Words &w, SimulationContext &ctx) {
    DumpArgs &d = ctx.dump_args;
    d.name = w.word();
    d.group = w.word();
    d.style = w.word();
    d.dump_step = w.count();
    d.file_name = w.word();
}

static void stmt_read_data(Words &w, SimulationContext &ctx);

struct Statement {
    const char *keyword;
    void (*read)(Words &, SimulationContext &);
};
static const Statement STATEMENTS[] = {
    {"timestep", stmt_timestep}, {"run", stmt_run},           {"velocity", stmt_velocity},     {"read_data", stmt_read_data},
    {"fix", stmt_fix},           {"pair_style", stmt_pair_style}, {"pair_coeff", stmt_pair_coeff}, {"dump", stmt_dump},
};

bool run_command(const std::string &command, const std::vector<std::string> &args, size_t line, SimulationContext &ctx) {
    for (const Statement &s : STATEMENTS)
        if (command == s.keyword) {
            Words w(args, line);
            s.read(w, ctx);
            return true;
        }
    return false;  // keywords are case-sensitive: "TimeStep" is unknown
}

// ---- LAMMPS data file -------------------------------------------------------------------------------
// Layout accepted (example/argon4000.txt, example/data_run.txt):
//     <N> atoms / <T> atom types / <lo> <hi> xlo xhi (ylo yhi, zlo zhi)          header fields, recognised by the word
//                                                                                 that FOLLOWS the numbers
//     Masses | PairCoeffs | Atoms | Velocities                                    section titles (first word of a line)
//     rows of the open section: Masses `type m`, PairCoeffs `type eps sigma [rc]`, Atoms `id type x y z`,
//     Velocities `id vx vy vz`; whatever follows the fields a row needs is ignored (trailing comments).
// Blank lines and lines starting with '#' are skipped.  Atom ids are 1-based and address the arrays directly.
class DataFile {
   public:
    explicit DataFile(double skin, int device) : mgr_(std::make_unique<LJCudaManager>(skin, device)) {}

    void read(std::istream &in) {
        std::string raw;
        for (size_t line = 0; std::getline(in, raw); ++line) {
            const std::string text = trim(raw);
            if (text.empty() || text[0] == '#') continue;
            Words w(split_whitespace(text), line);
            if (open_section(*w.peek())) continue;
            if (header_field(w)) continue;
            if (row_) (this->*row_)(w);
        }
    }

    // hand the result to the context: atoms + box, the potential if the file had a PairCoeffs section, and whether
    // velocities still have to be created (no Velocities section)
    void install(SimulationContext &ctx) {
        Atoms atoms;
        atoms.n_atoms = n_atoms_;
        atoms.type_ids = std::move(types_);
        atoms.masses = std::move(masses_);
        atoms.positions = std::move(x_);
        atoms.velocities = std::move(v_);
        atoms.forces.assign(3 * n_atoms_, 0.0);
        atoms.sim_box = SimulationBox::from_lammps_data(lo_[0], hi_[0], lo_[1], hi_[1], lo_[2], hi_[2], 0.0, 0.0, 0.0);
        ctx.atoms = std::move(atoms);
        if (!mgr_->is_empty()) ctx.mgr = std::move(mgr_);
        if (!ctx.starting_velocity) ctx.starting_velocity = StartVelocity{};
        ctx.starting_velocity->start_velocity = !has_velocities_;
    }

   private:
    using Row = void (DataFile::*)(Words &);

    bool open_section(const std::string &title) {
        static const struct {
            const char *title;
            Row row;
        } sections[] = {{"Masses", &DataFile::row_mass}, {"PairCoeffs", &DataFile::row_pair_coeff}, {"Atoms", &DataFile::row_atom},
                        {"Velocities", &DataFile::row_velocity}};
        for (const auto &s : sections)
            if (title == s.title) {
                row_ = s.row;
                if (s.row == &DataFile::row_velocity) has_velocities_ = true;
                return true;
            }
        return false;
    }

    // `<N> atoms`, `<T> atom types`, `<lo> <hi> xlo xhi`: the tag word sits right behind the numbers
    bool header_field(Words &w) {
        const std::string *tag1 = w.peek(1), *tag2 = w.peek(2);
        if (tag1 && *tag1 == "atoms") {
            n_atoms_ = w.count();
            types_.assign(n_atoms_, 0);
            x_.assign(3 * n_atoms_, 0.0);
            v_.assign(3 * n_atoms_, 0.0);
            return true;
        }
        if (tag1 && *tag1 == "atom") {
            masses_.assign(w.count(), 0.0);
            return true;
        }
        if (tag2) {
            static const char *const axis_tag[3] = {"xlo", "ylo", "zlo"};
            for (int d = 0; d < 3; ++d)
                if (*tag2 == axis_tag[d]) {
                    lo_[d] = w.real();
                    hi_[d] = w.real();
                    return true;
                }
        }
        return false;
    }

    void row_mass(Words &w) {
        const size_t type = w.count();
        const double m = w.real();
        if (type < 1 || type > masses_.size()) throw PisError("InvalidAtomType", "Atom type " + std::to_string(type) + " out of range");
        masses_[type - 1] = m;
    }

    // `type eps sigma [rc]` (rc defaults to 2.5 sigma).  A second word that does not read as a real selects the
    // two-type form `i j eps sigma [rc]` -- which an INTEGER j never does, because "2" reads as the real 2.0: the
    // reference mis-reads such a line as a like-pair entry with eps = j, and so does this reader.
    void row_pair_coeff(Words &w) {
        const int i = (int)w.count();
        int j = i;
        std::optional<double> eps = w.maybe_real();
        if (!eps) {
            j = (int)w.count();
            eps = w.real();
        }
        const double sigma = w.real();
        const double rc = w.maybe_real().value_or(2.5 * sigma);
        mgr_->insert({i, j}, LennardJones{*eps, sigma, rc, true});
    }

    size_t atom_index(Words &w) {
        const size_t id = w.count();
        if (id < 1 || id > n_atoms_)
            throw PisError("AtomCountMismatch", "Atom count mismatch: expected " + std::to_string(n_atoms_) + ", found " + std::to_string(id));
        return id - 1;
    }

    void row_atom(Words &w) {
        const size_t i = atom_index(w);
        types_[i] = (int32_t)w.count();
        for (int c = 0; c < 3; ++c) x_[3 * i + c] = w.real();
    }

    void row_velocity(Words &w) {
        const size_t i = atom_index(w);
        for (int c = 0; c < 3; ++c) v_[3 * i + c] = w.real();
    }

    Row row_ = nullptr;
    size_t n_atoms_ = 0;
    double lo_[3] = {0.0, 0.0, 0.0}, hi_[3] = {1.0, 1.0, 1.0};
    std::vector<double> masses_, x_, v_;
    std::vector<int32_t> types_;
    bool has_velocities_ = false;
    std::unique_ptr<LJCudaManager> mgr_;
};

// read_data <file>
static void stmt_read_data(Words &w, SimulationContext &ctx) {
    const std::string path = w.word();
    std::ifstream file(path);
    if (!file) throw PisError("InputFileError", "Failed to open input file '" + path + "': No such file or directory (os error 2))");
    DataFile data(ctx.skin, ctx.device);
    data.read(file);
    data.install(ctx);
}

// ---- DumpTraj (src/writers/dump_traj.rs) ---------------------------------------------------------------
DumpTraj::DumpTraj(const DumpArgs &args) : path_(args.file_name) {
    out_ = std::fopen(path_.c_str(), "w");
    if (!out_) throw PisError("DumpCreateError", "Failed to create trajectory file '" + path_ + "': " + std::strerror(errno));
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
    if (ctx.velocities_pending) {  // `velocity all create T seed`: generated where the state lives, host copy refreshed once
        mgr.start_velocities(ctx.starting_velocity->start_temperature.value_or(300.0), ctx.starting_velocity->seed.value_or(0));
        mgr.download(atoms, false, true, false);
        ctx.velocities_pending = false;
    }
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

// ---- System: script -> context -> run -----------------------------------------------------------------
// Script syntax: one statement per line, `keyword arg arg ...`; `#` starts a comment anywhere on a line; lines with fewer
// than two words are skipped; an unknown keyword is an error that names the line (1-based).
System &System::read() {
    std::ifstream file(infile_);
    if (!file) throw PisError("InputFileError", "Failed to open input file '" + infile_ + "': No such file or directory (os error 2))");
    std::string raw;
    for (size_t line = 1; std::getline(file, raw); ++line) {
        std::vector<std::string> words = split_whitespace(raw.substr(0, raw.find('#')));
        if (words.size() < 2) continue;
        const std::string keyword = words.front();
        words.erase(words.begin());
        if (!run_command(keyword, words, line, ctx))
            throw PisError("UnknownCommand", "Invalid command " + keyword + " found line: " + std::to_string(line));
    }
    return *this;
}

// `pair_style lj/cut <rc>` + `pair_coeff i j eps sigma [rc]` lines -> a manager that REPLACES one a data file installed.
// The table key is stored as written (i, j), while lookups use (min, max): `pair_coeff 2 1 ...` is never found
//  -- the device table reproduces that.
static std::unique_ptr<LJCudaManager> potential_from_script(const PotentialArgs &pa, double skin, int device) {
    Words style_line(pa.pair_style_args, pa.pair_style_line);
    const std::string style = style_line.word();
    if (style != "lj/cut") throw PisError("UnknownPairStyle", "Unknown pair style: '" + style + "'");
    const double global_cutoff = style_line.real();
    auto mgr = std::make_unique<LJCudaManager>(skin, device);
    for (size_t k = 0; k < pa.pair_coeff_args.size(); ++k) {
        Words w(pa.pair_coeff_args[k], pa.pair_coeff_lines[k]);
        const int i = (int)w.count(), j = (int)w.count();
        const double epsilon = w.real(), sigma = w.real();
        mgr->insert({i, j}, LennardJones{epsilon, sigma, w.maybe_real().value_or(global_cutoff), true});
    }
    return mgr;
}

System &System::contextualize() {
    if (ctx.atoms && ctx.atoms->n_atoms == 0) throw PisError("NoAtomsDefined", "No atoms defined in input file");
    // velocities: created unless a Velocities section supplied them (default 300 K, seed 0 without a velocity statement)
    if (ctx.atoms && ctx.starting_velocity && ctx.starting_velocity->start_velocity) {
        if (ctx.device_velocities) ctx.velocities_pending = true;  // Simulation::run: on the device, after the upload
        else ctx.atoms->start_velocities(ctx.starting_velocity->start_temperature.value_or(300.0), ctx.starting_velocity->seed.value_or(0));
    }
    if (ctx.potential_args) ctx.mgr = potential_from_script(*ctx.potential_args, ctx.skin, ctx.device);
    return *this;
}

void System::run(FILE *thermo_out) {
    if (!ctx.mgr) throw PisError("PotentialNotInitialized", "Potential manager not initialized - missing pair_style or pair_coeff commands");
    Simulation::run(*ctx.mgr, ctx, thermo_out);
}

// `--check`: the parsed context as one JSON object (no GPU needed), what tests/test_cli.py reads.
namespace {
struct Json {
    std::string s;
    bool first = true;
    void sep() {
        if (!first) s += ", ";
        first = false;
    }
    Json &key(const char *k) {
        sep();
        s += std::string("\"") + k + "\": ";
        first = true;
        return *this;
    }
    Json &raw(const std::string &v) {
        s += v;
        first = false;
        return *this;
    }
    Json &str(const std::string &v) { return raw("\"" + v + "\""); }
    Json &num(double v) { return raw(rust_display_f64(v)); }
    Json &open(char c) {
        s += c;
        first = true;
        return *this;
    }
    Json &close(char c) {
        s += c;
        first = false;
        return *this;
    }
    template <typename It, typename F>
    Json &list(It b, It e, F each) {
        open('[');
        for (; b != e; ++b) {
            sep();
            first = true;
            each(*b);
            first = false;
        }
        return close(']');
    }
};
}  // namespace

std::string System::describe() const {
    Json j;
    j.open('{');
    j.key("timestep").num(ctx.timestep);
    j.key("steps").raw(std::to_string(ctx.steps));
    j.key("dump").open('{');
    j.key("name").str(ctx.dump_args.name).key("group").str(ctx.dump_args.group).key("style").str(ctx.dump_args.style);
    j.key("dump_step").raw(std::to_string(ctx.dump_args.dump_step)).key("file_name").str(ctx.dump_args.file_name);
    j.close('}');
    j.key("ensemble").str(ctx.nh_chain_args ? (ctx.mtk_barostat_args ? "NPT" : "NVT") : "NVE");
    if (ctx.starting_velocity) {
        const StartVelocity &sv = *ctx.starting_velocity;
        j.key("velocity").open('{');
        j.key("group").str(sv.group).key("start_velocity").raw(sv.start_velocity ? "true" : "false");
        if (sv.start_temperature) j.key("temperature").num(*sv.start_temperature);
        if (sv.seed) j.key("seed").raw(std::to_string(*sv.seed));
        if (sv.dist) j.key("dist").str(*sv.dist);
        j.close('}');
    }
    if (ctx.atoms) {
        const Atoms &a = *ctx.atoms;
        const size_t head = std::min<size_t>(a.n_atoms, 4) * 3;
        auto num = [&](double v) { j.num(v); };
        j.key("atoms").open('{');
        j.key("n_atoms").raw(std::to_string(a.n_atoms)).key("n_types").raw(std::to_string(a.masses.size()));
        const double edges[3] = {a.sim_box.h[0], a.sim_box.h[4], a.sim_box.h[8]};
        j.key("box").list(edges, edges + 3, num);
        j.key("first_positions").list(a.positions.begin(), a.positions.begin() + (std::ptrdiff_t)head, num);
        j.key("first_velocities").list(a.velocities.begin(), a.velocities.begin() + (std::ptrdiff_t)head, num);
        j.key("masses").list(a.masses.begin(), a.masses.end(), num);
        j.key("types_head").list(a.type_ids.begin(), a.type_ids.begin() + (std::ptrdiff_t)std::min<size_t>(a.n_atoms, 8),
                                 [&](int32_t t) { j.raw(std::to_string(t)); });
        j.close('}');
    }
    j.key("potential");
    if (ctx.mgr) {
        j.open('{');
        j.key("max_rcut").num(ctx.mgr->max_rcut());
        j.key("pairs").list(ctx.mgr->table.begin(), ctx.mgr->table.end(), [&](const std::pair<const std::pair<int, int>, LennardJones> &kv) {
            j.open('{');
            j.key("i").raw(std::to_string(kv.first.first)).key("j").raw(std::to_string(kv.first.second));
            j.key("epsilon").num(kv.second.epsilon).key("sigma").num(kv.second.sigma).key("rcut").num(kv.second.rcut);
            j.close('}');
        });
        j.close('}');
    } else {
        j.raw("null");
    }
    j.close('}');
    return j.s;
}

}  // namespace pis
