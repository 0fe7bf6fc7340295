"""profiles/r2/parity_log.jsonl (written by tests/parity.py during `pytest -m gpu` on the B200 box) -> profiles/r2_parity.md."""
import collections
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2/parity_log.jsonl"
rows = [json.loads(l) for l in open(src)]
par = [r for r in rows if r["kind"] == "parity"]
wt = [r for r in par if "truth_viol" in r]
out = ["# Round 2: achieved parity, every `assert_parity` call of one green `pytest -m gpu` run on a B200", "",
       f"{len(par)} assertions, {sum(r['n'] for r in par):,} output elements; tolerance rtol=1e-2 / atol=1e-3 (north_star), hard bound "
       "rtol=2e-2 / atol=8e-3.  Source: `profiles/r2/parity_log.jsonl` (one JSON line per assertion).", "",
       "## Against the oracle (all assertions)", "",
       f"* elements outside the tolerance: {sum(r['viol'] for r in par):,} of {sum(r['n'] for r in par):,} "
       f"({sum(r['viol'] for r in par) / sum(r['n'] for r in par):.2e}); outside the hard bound: {sum(r['hard'] for r in par)}",
       f"* worst single assertion: {max(par, key=lambda r: r['viol'] / r['n'])['viol']} of {max(par, key=lambda r: r['viol'] / r['n'])['n']} "
       f"elements ({max(r['viol'] / r['n'] for r in par):.2e}) — `{max(par, key=lambda r: r['viol'] / r['n'])['what']}`",
       "", "## Where FlashAttention-2 ran on the same inputs (the relative gate)", "",
       f"{len(wt)} assertions, {sum(r['n'] for r in wt):,} elements.  Counts of elements outside rtol=1e-2 / atol=1e-3:", "",
       "| measured against | ours | FlashAttention-2 (installed 2.8.3) |", "|---|---:|---:|",
       f"| the oracle (fp32 softmax, P rounded against the running max like FA2) | {sum(r['viol'] for r in wt):,} | {sum(r['fa2_viol'] for r in wt):,} |",
       f"| exact fp64 attention | {sum(r['truth_viol'] for r in wt):,} | {sum(r['fa2_truth_viol'] for r in wt):,} |",
       f"| hard bound, exact fp64 | {sum(r['truth_hard'] for r in wt)} | {sum(r['fa2_truth_hard'] for r in wt)} |", "",
       "Per test (sum over its chunks), sorted by our excess over FA2 against exact math:", "",
       "| test | elements | ours vs oracle | FA2 vs oracle | ours vs exact | FA2 vs exact |", "|---|---:|---:|---:|---:|---:|"]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in wt:
    a = agg[r["test"].split(" ")[0].split("::")[-1]]
    for i, k in enumerate(("n", "viol", "fa2_viol", "truth_viol", "fa2_truth_viol")):
        a[i] += r[k]
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][3] - kv[1][4]))[:25]:
    out.append(f"| `{k}` | {a[0]:,} | {a[1]} | {a[2]} | {a[3]} | {a[4]} |")
worst = max(wt, key=lambda r: (r["truth_viol"] - r["fa2_truth_viol"]) / r["n"])
out += ["", f"Largest excess of one assertion: {(worst['truth_viol'] - worst['fa2_truth_viol']) / worst['n']:.2e} of the elements "
        f"({worst['truth_viol']} vs {worst['fa2_truth_viol']} of {worst['n']:,}; `{worst['test'].split('::')[-1].split(' ')[0]}`, {worst['what']}) "
        "— the sharp-softmax stress cases (logit std 6-8) on the tcgen05 kernel, whose lazy softmax reference leaves the RMS "
        "error ~1.25x FA2's there; `tests/parity.py: REL_EPS` = 2e-3 is set from this measurement.", ""]
other = [r for r in rows if r["kind"] != "parity"]
if other:
    out += ["## Benchmarked shape (32,768-token chunk over 98,304 cached tokens), sampled rows vs exact fp64", "",
            "| n_full | violations ours | violations FA2 | RMS err ours | RMS err FA2 | max err ours | max err FA2 |", "|---:|---:|---:|---:|---:|---:|---:|"]
    for r in other:
        out.append(f"| {r['n_full']} | {r['viol_ours']:.1e} | {r['viol_fa2']:.1e} | {r['rms_ours']:.2e} | {r['rms_fa2']:.2e} | "
                   f"{r['max_ours']:.2e} | {r['max_fa2']:.2e} |")
open("profiles/r2_parity.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:30]))
