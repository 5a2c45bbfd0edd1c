"""Regenerates oracle_vectors.json: full-precision outputs of the CPU oracle on the shipped parameter.txt (r_norm,
GMRES counts, residual history, the four error columns, lambda) for the linear mortar (cycles 0-2) and the quadratic
mortar under both tie rules (cycles 0-2).  The file pins the oracle against drift (tests/test_oracle_vectors.py) and
gives the GPU parity tests numbers that do not need the oracle at run time.  Run from the repository root:
    python tests/golden/make_oracle_vectors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmmfe_oracle as mo  # noqa: E402

COLS = ["velocity_l2l2", "pressure_dg", "pressure_l2l2", "lambda_l2"]


def case(mortar_degree, tie_rules, cycles=3):
    prm = mo.parse_parameter_file(os.path.join(ROOT, "parameter.txt"))
    prm.mortar_degree, prm.num_refinement = mortar_degree, cycles
    out = []
    for res in mo.run(prm, tie_rules=tie_rules):
        out.append({"cycle": res.cycle, "gmres_iterations": res.gmres_iterations, "r_norm": float(res.r_norm),
                    "errors": {c: float(res.errors[c]) for c in COLS},
                    "residuals": [float(x) for x in res.residuals], "lambda": [float(x) for x in res.lam]})
    return out


if __name__ == "__main__":
    data = {"source": "oracle/mmmfe_oracle.py on parameter.txt (tolerance 1e-6), written by tests/golden/make_oracle_vectors.py",
            "linear": case(1, None), "quadratic_lower": case(2, None), "quadratic_dealii91": case(2, "dealii91")}
    with open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json"), "w") as fh:
        json.dump(data, fh)
    print({k: [(r["gmres_iterations"], len(r["lambda"])) for r in v] for k, v in data.items() if k != "source"})
