"""Turn the ncu artefacts in gpurun_out/ (written by profiles/capture.sh on the B200 box) into the small,
tracked summaries under profiles/.  Run here (no GPU needed): python profiles/summarize.py r1"""
import collections
import csv
import io
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
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


with open(f"{OUT}/{tag}_decode.md", "w") as f:
    f.write(f"# {tag}: decode @1M ctx, Llama-3-8B pattern 0.5 (command: profiles/capture.sh)\n\n")
    f.write(launch_table(f"{SRC}/launches_decode.csv", "Launch list of ONE timed decode step (nvtx range timed_decode)"))
    f.write("\n")
    f.write(kernel_section(f"{SRC}/prof_decode.ncu-rep", "`ncu --set full` of duo_attn_mma_kernel (3 launches)",
                           "traffic = dram__bytes_read.sum + dram__bytes_write.sum; algorithmic bytes of a launch = "
                           "n_full * (N+1) * 512 B (+ streaming heads, negligible)."))
with open(f"{OUT}/{tag}_prefill.md", "w") as f:
    f.write(f"# {tag}: prefill 128K in 32K chunks, Llama-3-8B pattern 0.5 (command: profiles/capture.sh)\n\n")
    f.write(launch_table(f"{SRC}/launches_prefill.csv", "Launch list of ONE timed 128K prefill (nvtx range timed_prefill)"))
    f.write("\n")
    f.write(kernel_section(f"{SRC}/prof_prefill.ncu-rep", "`ncu --set full` of duo_attn_tc_kernel (2 launches, 4th chunk)",
                           "Tensor-pipe utilisation = sm__ops_path_tensor_op_utchmma_* / TriageCompute.sm__pipe_tensor_cycles_active."))
print("wrote", f"{OUT}/{tag}_decode.md", f"{OUT}/{tag}_prefill.md")
