"""Diffs the output of java/.../Harness.java (Jstacs itself) against tests/golden/hmm_golden.json (the oracle).

usage: python java/compare_with_golden.py jstacs_golden.json tests/golden/hmm_golden.json
Gates (BASELINE.json north_star): Viterbi paths bit-exact, Viterbi scores bit-exact, log-likelihoods 1e-9 relative,
objectives 1e-9 relative (progress lines print full doubles).  Parameters after training are compared through the
objective of the last iteration (layouts differ: flat Java layout vs jh_model_desc layout)."""
import json
import sys


def fx(v):
    return [float.fromhex(x) for x in v]


def main():
    java = json.load(open(sys.argv[1]))["cases"]
    gold = json.load(open(sys.argv[2]))["cases"]
    bad = 0
    for name, g in gold.items():
        j = java.get(name)
        if j is None:
            print(f"{name}: missing in the Java output")
            bad += 1
            continue
        ok_paths = j["viterbi_paths"] == g["viterbi_paths"]
        ok_scores = fx(j["viterbi_scores"]) == fx(g["viterbi_scores"])
        rel = lambda a, b: max((abs(x - y) / max(abs(y), 1e-300) for x, y in zip(a, b)), default=0.0)
        r_ll = rel(fx(j["loglik"]), fx(g["loglik"]))
        r_obj = rel(fx(j["objective_4_iterations"]), fx(g["objective_4_iterations"]))
        ok = ok_paths and ok_scores and r_ll <= 1e-9 and r_obj <= 1e-9
        bad += not ok
        print(f"{name:18s} paths {'==' if ok_paths else '!='}  scores {'==' if ok_scores else '!='}  loglik rel {r_ll:.2e}  objective rel {r_obj:.2e}  {'OK' if ok else 'MISMATCH'}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
