"""Pin the oracle against EXACT ARITHMETIC instead of against itself.

  python tests/golden/make_exact.py        ->  tests/golden/exact_small.json

The reference ships no golden vectors for the hot path and cannot be built here (Rust), so the CPU oracle
(oracle/pis_oracle.c) would otherwise only be pinned by fixtures it produced itself.  This script evaluates the
FORMULAS of the reference -- not any implementation of them -- with 60-digit mpmath arithmetic on the committed
108-atom fixture (tests/golden/oracle_small.json: positions, velocities, box, rc, skin), by a plain O(N^2) loop
over all pairs:

  minimum image   d - L * round(d / L)                                   src/simulation_box.rs:17-27 (orthorhombic)
  cutoff          |rij| > rcut -> skip (inclusive)                       src/potentials/lennard_jones.rs:224
  pair            u = 4 eps (s^12 - s^6) - u_cut,  s^2 = sigma^2 / r^2,
                  f = 24 eps (2 s^12 - s^6) / r^2 * rij  (on j; -f on i)  src/potentials/lennard_jones.rs:33-55
  list            j in list(i)  <=>  |rij| <= rcut + skin                src/potentials/lennard_jones.rs:393-405
  one NVE step    x += v dt + (F/m) 0.5 dt^2; wrap; F' = F(x');
                  v += (F/m + F'/m) 0.5 dt                               src/potentials/potential.rs:15-33
  observables     KE = sum 0.5 m v^2, T = 2 KE / (3 N kB),
                  tr(X F^T) = sum r_i . F_i                              src/atoms/properties.rs:17-65

The inputs are the exact binary values of the fixture's doubles; results are rounded once, to the nearest double.
A 64-bit restatement that follows the reference's operation order agrees with these numbers to a few ulp of the
largest term; tests/test_oracle.py::test_oracle_matches_exact_arithmetic holds the oracle to 1e-13.
The margins (how far the closest pair sits from each cutoff) are stored so the test can assert that no in/out
decision is within rounding distance of a threshold, i.e. that the exact sets are unambiguous in f64.
"""
import json
import os

from mpmath import mp, mpf, nint, sqrt

HERE = os.path.dirname(os.path.abspath(__file__))
KB = mpf("0.0083144621")  # src/constants.rs:3


def main():
    mp.dps = 60
    with open(os.path.join(HERE, "oracle_small.json")) as fh:
        g = json.load(fh)
    L, rc, skin, dt = mpf(g["L"]), mpf(g["rc"]), mpf(g["skin"]), mpf(g["dt"])
    eps, sigma, mass = mpf(g["eps"]), mpf(g["sigma"]), mpf(g["mass"])
    x = [[mpf(c) for c in r] for r in g["positions"]]
    v = [[mpf(c) for c in r] for r in g["velocities"]]
    n = len(x)
    sr2 = (sigma / rc) ** 2
    ucut = 4 * eps * (sr2 ** 6 - sr2 ** 3)

    def evaluate(pos):
        pe = mpf(0)
        frc = [[mpf(0)] * 3 for _ in range(n)]
        rows = [[] for _ in range(n)]
        margin_rc, margin_list = mpf("inf"), mpf("inf")
        for i in range(n):
            for j in range(i + 1, n):
                d = [pos[j][k] - pos[i][k] for k in range(3)]
                d = [c - L * nint(c / L) for c in d]
                r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
                r = sqrt(r2)
                margin_rc = min(margin_rc, abs(r - rc))
                margin_list = min(margin_list, abs(r - (rc + skin)))
                if r <= rc + skin:
                    rows[i].append(j)
                    rows[j].append(i)
                if r > rc:
                    continue
                s2 = sigma * sigma / r2
                s6 = s2 ** 3
                s12 = s6 * s6
                pe += 4 * eps * (s12 - s6) - ucut
                fs = 24 * eps * (2 * s12 - s6) / r2
                for k in range(3):
                    frc[i][k] -= fs * d[k]
                    frc[j][k] += fs * d[k]
        return pe, frc, rows, margin_rc, margin_list

    pe0, f0, rows0, m_rc0, m_list0 = evaluate(x)
    # one verlet_step_nve
    x1 = [[x[i][k] + v[i][k] * dt + (f0[i][k] / mass) * mpf("0.5") * dt * dt for k in range(3)] for i in range(n)]
    x1 = [[c - L * mp.floor(c / L) for c in r] for r in x1]
    pe1, f1, _, m_rc1, _ = evaluate(x1)
    v1 = [[v[i][k] + (f0[i][k] / mass + f1[i][k] / mass) * mpf("0.5") * dt for k in range(3)] for i in range(n)]
    ke1 = sum(mpf("0.5") * mass * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) for r in v1)
    t1 = 2 * ke1 / (3 * n * KB)
    vir1 = sum(x1[i][k] * f1[i][k] for i in range(n) for k in range(3))
    vol = L ** 3
    p1 = (2 * ke1 + vir1) / (3 * vol)

    def fl(a):
        return float(a)

    out = {
        "source": "60-digit mpmath evaluation of the reference's formulas on tests/golden/oracle_small.json (make_exact.py)",
        "n": n, "pe0": fl(pe0), "forces0": [[fl(c) for c in r] for r in f0],
        "neighbours_skin": [sorted(r) for r in rows0],
        "margin_rc0": fl(m_rc0), "margin_list0": fl(m_list0), "margin_rc1": fl(m_rc1),
        "positions1": [[fl(c) for c in r] for r in x1], "velocities1": [[fl(c) for c in r] for r in v1],
        "forces1": [[fl(c) for c in r] for r in f1], "pe1": fl(pe1), "ke1": fl(ke1), "temperature1": fl(t1),
        "virial_trace1": fl(vir1), "pressure1": fl(p1),
        "pe0_digits": mp.nstr(pe0, 30), "pe1_digits": mp.nstr(pe1, 30), "ke1_digits": mp.nstr(ke1, 30),
    }
    with open(os.path.join(HERE, "exact_small.json"), "w") as fh:
        json.dump(out, fh)
    print("exact_small.json written: pe0 =", out["pe0_digits"], " margins:", out["margin_rc0"], out["margin_list0"], out["margin_rc1"])


if __name__ == "__main__":
    main()
