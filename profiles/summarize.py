"""Turn the ncu artefacts in gpurun_out/ (written by profiles/capture.sh on the B200 box) into the small,
tracked summaries under profiles/.  Run here (no GPU needed): python profiles/summarize.py r1"""
import collections
import csv
import io
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT = "profiles"
SRC = "gpurun_out"


def launches(fn):
    lines = open(fn).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), float(r["Metric Value"].replace(",", ""))))
    return rows


def launch_table(fn, title):
    rows = launches(fn)
    tot = sum(v for _, v in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[re.sub(r"<.*", "", k)[:80]][0] += 1
        agg[re.sub(r"<.*", "", k)[:80]][1] += v
    out = [f"### {title}", "",
           f"{len(rows)} launches, {tot/1e6:.3f} ms total device time (ncu: cold cache, serialised — compare SHARES, not absolutes)",
           "", "| share | launches | avg (us) | kernel |", "|---:|---:|---:|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f"| {v/tot*100:.2f}% | {n} | {v/n/1e3:.1f} | `{k}` |")
    return "\n".join(out) + "\n"


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max.per_second",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for row in rd[2:]:
        d = {"Kernel Name": row[idx["Kernel Name"]]}
        for w in WANT:
            if w in idx:
                d[w] = (row[idx[w]], units[idx[w]])
        out.append(d)
    return out


def stalls(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    first = []
    for r in rows[2:]:
        if r[0] == "Kernel Name":
            break
        first.append(r)
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]]) for r in first)
    agg = sorted(((h, sum(int(r[idx[h]]) for r in first)) for h in cols), key=lambda kv: -kv[1])
    hot = sorted(first, key=lambda r: -int(r[idx["# Samples"]]))[:14]
    out = ["| stall reason | samples | share |", "|---|---:|---:|"]
    out += [f"| {h} | {v} | {v/tot*100:.1f}% |" for h, v in agg[:8]]
    out += ["", "| share of samples | executed | SASS |", "|---:|---:|---|"]
    out += [f"| {int(r[idx['# Samples']])/tot*100:.2f}% | {r[idx['Instructions Executed']]} | `{r[1].strip()[:90]}` |" for r in hot]
    return "\n".join(out) + "\n"


def kernel_section(rep, title, note):
    out = [f"### {title}", "", note, ""]
    for d in raw_metrics(rep):
        out.append(f"`{d['Kernel Name'][:110]}`\n")
        out.append("| metric | value | unit |\n|---|---:|---|")
        for w in WANT:
            if w in d:
                out.append(f"| {w} | {d[w][0]} | {d[w][1]} |")
        out.append("")
    out.append("Warp-stall sampling, first captured launch:\n")
    out.append(stalls(rep))
    return "\n".join(out)


import json
import os

CTX = 1048576


def traffic_entry(rep, row_bytes, source):
    """DRAM bytes of the first captured launch + the algorithmic bytes of that same launch (its number of retrieval
    heads is recovered from the traffic itself: n_full = round(read / ((ctx + 1) * row_bytes)))."""
    d = raw_metrics(rep)[0]

    def val(name):
        v, unit = d[name]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    n_full = max(1, round(rd / ((CTX + 1) * row_bytes)))
    alg = n_full * (CTX + 1) * row_bytes + (8 - n_full) * 321 * row_bytes
    return {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": alg, "n_full_of_launch": n_full,
            "source": source}


traffic = {}
if os.path.exists(f"{SRC}/prof_decode.ncu-rep"):
    with open(f"{OUT}/{tag}_decode.md", "w") as f:
        f.write(f"# {tag}: decode @1M ctx, Llama-3-8B pattern 0.5 (command: profiles/capture.sh)\n\n")
        f.write(launch_table(f"{SRC}/launches_decode.csv", "Launch list of ONE timed decode step (nvtx range timed_decode)"))
        f.write("\n")
        f.write(kernel_section(f"{SRC}/prof_decode.ncu-rep", "`ncu --set full` of duo_attn_mma_kernel (fused decode step)",
                               "traffic = dram__bytes_read.sum + dram__bytes_write.sum; algorithmic bytes of a launch = "
                               "n_full * (N+1) * 512 B (+ streaming heads, negligible)."))
    traffic["decode"] = traffic_entry(f"{SRC}/prof_decode.ncu-rep", 512, f"profiles/{tag}_decode.md")
if os.path.exists(f"{SRC}/prof_prefill.ncu-rep"):
    with open(f"{OUT}/{tag}_prefill.md", "w") as f:
        f.write(f"# {tag}: prefill 128K in 32K chunks, Llama-3-8B pattern 0.5 (command: profiles/capture.sh)\n\n")
        f.write(launch_table(f"{SRC}/launches_prefill.csv", "Launch list of ONE timed 128K prefill (nvtx range timed_prefill)"))
        f.write("\n")
        f.write(kernel_section(f"{SRC}/prof_prefill.ncu-rep", "`ncu --set full` of duo_attn_tc_kernel (last 32K chunk over 96K past, n_full = 4)",
                               "Tensor-pipe utilisation = sm__ops_path_tensor_op_utchmma_* / TriageCompute.sm__pipe_tensor_cycles_active."))
    d = raw_metrics(f"{SRC}/prof_prefill.ncu-rep")[0]
    traffic["prefill"] = {"dram_bytes": sum(float(d[k][0].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d[k][1], 1)
                                            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")),
                          "source": f"profiles/{tag}_prefill.md"}
if os.path.exists(f"{SRC}/prof_int4_dec8.ncu-rep"):
    with open(f"{OUT}/{tag}_int4.md", "w") as f:
        f.write(f"# {tag}: INT4-KV decode @1M ctx, Llama-3-8B pattern 0.5 (command: profiles/capture_int4.sh)\n\n")
        if os.path.exists(f"{SRC}/launches_int4.csv"):
            f.write(launch_table(f"{SRC}/launches_int4.csv", "Launch list of ONE timed INT4 decode step"))
            f.write("\n")
        f.write(kernel_section(f"{SRC}/prof_int4_dec8.ncu-rep", "`ncu --set full` of duo_attn_int4_dec8_kernel",
                               "algorithmic bytes of a launch = n_full * (N+1) * 2 * 68 B (64 B codes + fp16 scale + fp16 zero per "
                               "row, K and V) + streaming heads (negligible)."))
    traffic["decode_int4"] = traffic_entry(f"{SRC}/prof_int4_dec8.ncu-rep", 136, f"profiles/{tag}_int4.md")
if traffic:
    old = {}
    if os.path.exists(f"{OUT}/{tag}_traffic.json"):
        old = json.load(open(f"{OUT}/{tag}_traffic.json"))
    old.update(traffic)
    json.dump(old, open(f"{OUT}/{tag}_traffic.json", "w"), indent=1)
print("wrote", [k for k in traffic], "->", f"{OUT}/{tag}_*.md / {tag}_traffic.json")
