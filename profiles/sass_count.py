"""Static issue-slot census of a kernel's main loop from its SASS (no GPU needed).

    python profiles/sass_count.py duo_attention_b200/csrc/attn_int4.o duo_attn_int4_dec8_kernel

Finds the innermost backward branch whose body contains HMMA (the K/V tile loop), prints the opcode histogram of the
loop body and of the *straight-line hot path* (forward branches over BSSY-guarded cold regions are followed as taken
when the guarded region contains a CALL, a second-level BSSY, or more than `--cold` instructions — i.e. the mask /
rescale / boundary-loader paths that interior tiles skip)."""
import argparse
import collections
import re
import subprocess


def load(obj, pattern):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn = [m for m in re.findall(r"Function : (\S+)", names) if pattern in m]
    if not fn:
        raise SystemExit(f"no function matching {pattern!r}")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], obj], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return fn[0], ins


def opcode(text):
    text = re.sub(r"^@!?U?P\w+\s+", "", text)
    return text.split()[0].split(".")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--cold", type=int, default=60)
    a = ap.parse_args()
    name, ins = load(a.obj, a.kernel)
    addr = {ad: i for i, (ad, _) in enumerate(ins)}
    loops = []
    for i, (ad, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < ad and int(m.group(1), 16) in addr:
            lo = addr[int(m.group(1), 16)]
            body = ins[lo: i + 1]
            if any("HMMA" in x for _, x in body):
                loops.append((len(body), lo, i))
    if not loops:
        raise SystemExit("no HMMA loop found")
    _, lo, hi = min(loops)
    body = ins[lo: hi + 1]
    print(f"{name}\nloop body: {len(body)} instructions [{ins[lo][0]:#x} .. {ins[hi][0]:#x}]")
    # hot path: walk, skipping BSSY-guarded regions that look cold
    hot = []
    i = lo
    while i <= hi:
        ad, t = ins[i]
        m = re.search(r"BSSY\S*\s+B\d+,\s+(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr:
            end = addr[int(m.group(1), 16)]
            region = ins[i + 1: end]
            cold = (len(region) > a.cold) or any("CALL" in x or "BSSY" in x for _, x in region)
            if cold and end <= hi + 1:
                # a guarded region is entered through a conditional branch placed before its body; the instructions
                # up to that branch are executed, the rest is skipped by interior tiles
                j = i + 1
                while j < end and not re.search(r"\bBRA\b", ins[j][1]):
                    j += 1
                hot.extend(ins[i: j + 1])
                i = end
                continue
        # unconditional-ish forward jumps over a cold block (e.g. the rescale block behind VOTE)
        m = re.search(r"^@!?P\d+\s+BRA\s+(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr and int(m.group(1), 16) > ad:
            end = addr[int(m.group(1), 16)]
            if end - i > a.cold and end <= hi + 1 and i > lo and "VOTE" in " ".join(x for _, x in ins[max(lo, i - 3): i]):
                hot.append(ins[i])
                i = end
                continue
        hot.append(ins[i])
        i += 1
    for title, seq in (("whole loop body", body), ("hot path (interior tile, running max unchanged)", hot)):
        c = collections.Counter(opcode(t) for _, t in seq)
        tot = sum(c.values())
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
        print(f"{title}: {tot} issue slots\n   {top}")


if __name__ == "__main__":
    main()
