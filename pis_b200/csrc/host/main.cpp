// main.cpp -- `pis_b200_cli`, the host CLI mirroring the reference binary (src/main.rs:50-62,
// src/args_parser.rs:5-10): `-i/--infile <script>` (default input.pis), `Error: <msg>` + exit 1 on failure.
// Extra flags the reference does not have (kept out of the script so inputs stay reference-compatible):
//   --skin <A>    Verlet skin distance handed to the CUDA manager (default: 0.3 x the largest sigma, or 1.0)
//   --device <n>  CUDA device ordinal
//   --check       parse + contextualize only, print the context as JSON (no GPU needed)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pis_host.hpp"

int main(int argc, char **argv) {
    std::string infile = "input.pis";
    double skin = -1.0;
    int device = 0;
    bool check = false;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        auto next = [&]() -> const char * {
            if (k + 1 >= argc) {
                std::fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", a.c_str());
                std::exit(2);
            }
            return argv[++k];
        };
        if (a == "-i" || a == "--infile") infile = next();
        else if (a.rfind("--infile=", 0) == 0) infile = a.substr(9);
        else if (a == "--skin") skin = std::atof(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--check") check = true;
        else if (a == "-h" || a == "--help") {
            std::printf("Usage: pis_b200_cli [-i|--infile <FILE>] [--skin <A>] [--device <N>] [--check]\n");
            return 0;
        } else {
            std::fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str());
            return 2;
        }
    }
    try {
        pis::System sys(infile);
        sys.ctx.device = device;
        sys.ctx.skin = skin < 0.0 ? 1.0 : skin;
        sys.ctx.device_velocities = !check;  // a run creates velocities on the device; --check needs no GPU
        sys.read().contextualize();
        if (skin < 0.0 && sys.ctx.mgr) {  // default skin: 0.3 x the largest sigma
            double smax = 0.0;
            for (auto &kv : sys.ctx.mgr->table) smax = std::max(smax, kv.second.sigma);
            if (smax > 0.0) {
                auto fresh = std::make_unique<pis::LJCudaManager>(0.3 * smax, device);
                for (auto &kv : sys.ctx.mgr->table) fresh->insert(kv.first, kv.second);
                sys.ctx.mgr = std::move(fresh);
            }
        }
        if (check) {
            std::printf("%s\n", sys.describe().c_str());
            return 0;
        }
        sys.run(stdout);
    } catch (const pis::PisError &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
