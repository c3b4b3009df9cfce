#!/usr/bin/env python
"""Generates tests/golden/*.json from the pure-Python restatement oracle/pyref.py (NOT from a run of the Rust
reference: no cargo in this image).  Deterministic; re-run to regenerate:
    python tests/golden/make_golden.py
Each file pins complete proofs (every round message, challenge and final value) for seeded inputs, so the C
oracle, the host library code and the CUDA path are all checked against the same stored bytes."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref as R

FIELDS = {0: R.BN254_FR, 1: R.BN254_FQ, 2: R.BLS12_381_FR}


def hx(v):
    return [hex(int(x)) for x in v]


def main():
    out = {"plain": [], "composed": [], "gkr": []}
    for fid, p in FIELDS.items():
        rng = random.Random(1000 + fid)
        for n in (1, 3, 6):
            tab = [rng.randrange(p) for _ in range(1 << n)]
            pr = R.prove(R.MultilinearPoly(tab, p))
            out["plain"].append({"field": fid, "n": n, "table": hx(tab), "claimed_sum": hex(pr.claimed_sum),
                                 "msgs": [hx(m) for m in pr.proof_polynomials], "challenges": hx(pr.challenges)})
        for mode, P, D, n in (("compat", 2, 2, 4), ("compat", 2, 3, 3), ("full", 1, 2, 5), ("full", 2, 3, 4), ("full", 1, 4, 3)):
            tabs = [[rng.randrange(p) for _ in range(1 << n)] for _ in range(P * D)]
            sp = R.SumPoly([R.ProductPoly(tabs[q * D:(q + 1) * D], p) for q in range(P)])
            pr = R.gkr_prove(0, sp, R.Transcript(p), mode)
            out["composed"].append({"field": fid, "mode": mode, "P": P, "D": D, "n": n, "tables": [hx(t) for t in tabs],
                                    "coeffs": [hx(c) for c in pr.proof_polynomials], "challenges": hx(pr.random_challenges)})
        for n_layers, out_gates in ((1, 1), (3, 1), (4, 2)):
            gates = [out_gates << (n_layers - 1 - l) for l in range(n_layers)]
            ops = [[rng.randrange(2) for _ in range(g)] for g in gates]
            inputs = [rng.randrange(p) for _ in range(2 * gates[0])]
            circ = R.Circuit([list(o) for o in ops])
            pr = R.gkr_protocol_prove_dense(circ, inputs, p)  # the reference's dense construction
            out["gkr"].append({"field": fid, "gates": gates, "ops": ops, "inputs": hx(inputs), "output_poly": hx(pr.output_poly),
                               "proof_polynomials": [[hx(c) for c in layer] for layer in pr.proof_polynomials],
                               "claimed_evaluations": [hx(ce) for ce in pr.claimed_evaluations],
                               "final_openings": hx(pr.final_openings)})
    with open(os.path.join(HERE, "proofs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(HERE, "proofs.json"))
    # general wiring (extension): proofs from the DENSE construction with general indices
    wired = []
    for fid, p in FIELDS.items():
        rng = random.Random(2000 + fid)
        for nin, gates in ((2, [1]), (4, [4, 2]), (8, [4, 8, 4]), (4, [8, 4, 1])):
            spec, w = [], nin
            for G in gates:
                spec.append({"ops": [rng.randrange(2) for _ in range(G)], "in1": [rng.randrange(w) for _ in range(G)],
                             "in2": [rng.randrange(w) for _ in range(G)]})
                w = G
            inputs = [rng.randrange(p) for _ in range(nin)]
            ws, w = [], nin
            for l in spec:
                ws.append(R.WiredLayer(l["ops"], l["in1"], l["in2"], w))
                w = len(l["ops"])
            pr = R.wired_prove_dense(R.WiredCircuit(ws), inputs, p)
            wired.append({"field": fid, "n_inputs": nin, "layers": spec, "inputs": hx(inputs), "output_poly": hx(pr.output_poly),
                          "proof_polynomials": [[hx(c) for c in layer] for layer in pr.proof_polynomials],
                          "claimed_evaluations": [hx(ce) for ce in pr.claimed_evaluations],
                          "final_openings": hx(pr.final_openings)})
    with open(os.path.join(HERE, "wired_proofs.json"), "w") as f:
        json.dump({"wired": wired}, f, indent=0)
    print("wrote", os.path.join(HERE, "wired_proofs.json"))


if __name__ == "__main__":
    main()
