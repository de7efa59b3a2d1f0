"""Extract the golden vectors the reference ships for the PNP hot path into small JSON fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Source: src/simulation_states/equilibrium/cyl/*_p0.dat of jurgispods/dune-ax1 -- converged
equilibrium states written by Ax1SimulationState (dune/ax1/common/ax1_simulationstate.hh:30-57):
grid vectors x,y, serialized channel states, uold, unew (vertex-blocked [Na,K,Cl,pot], x-fastest).
Only DATA is copied (no reference source code).
"""
import json
import os
import sys

REF = "/root/reference/src/simulation_states/equilibrium/cyl"
OUT = os.path.dirname(os.path.abspath(__file__))
FILES = ["yasp_square10mm_p0.dat", "yasp_square1mm_p0.dat", "yasp_square10mm_myelin_p0.dat"]


def parse(path):
    tok = open(path).read().split()
    i = 0
    out = {}
    vec_names = []
    while i < len(tok):
        key = tok[i]
        if key in ("#elements", "#vectors"):
            out[key.strip("#")] = int(tok[i + 1]); i += 2
        elif key in ("time", "dt", "timestep", "mpi_rank", "mpi_size"):
            out[key] = float(tok[i + 1]); i += 2
        else:
            n = int(tok[i + 1])
            # keep the decimal strings: they round-trip the 17 significant digits exactly
            out[key] = [float(t) for t in tok[i + 2:i + 2 + n]]
            vec_names.append(key)
            i += 2 + n
    out["vector_order"] = vec_names
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted; fixtures are committed, nothing to do")
    for f in FILES:
        d = parse(os.path.join(REF, f))
        dst = os.path.join(OUT, f.replace("_p0.dat", ".json"))
        with open(dst, "w") as fh:
            json.dump(d, fh, separators=(",", ":"))
        print(dst, {k: (len(v) if isinstance(v, list) else v) for k, v in d.items()})
        # the state files themselves are input data of a run (loadState): verbatim copies for the package
        import shutil
        data = os.path.join(os.path.dirname(os.path.dirname(OUT)), "dune_ax1_b200", "data", f)
        shutil.copyfile(os.path.join(REF, f), data)
        os.chmod(data, 0o644)


if __name__ == "__main__":
    main()
