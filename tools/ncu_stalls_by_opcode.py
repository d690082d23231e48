#!/usr/bin/env python
"""Per-SASS-opcode stall table from the source page of an .ncu-rep captured with --import-source on (read here, on the CPU box).

    python tools/ncu_stalls_by_opcode.py gpurun_out/prof_wino_log_prob.ncu-rep profiles/r06_chain_winograd_stalls_by_opcode.txt "header text"
"""
import collections
import csv
import re
import subprocess
import sys


def main(rep, out, header=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    col = {h: i for i, h in enumerate(hdr)}
    by_op = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
    by_reason = collections.Counter()
    per_ins = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]]
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src.split() else "?"
        st = {k: int(r[col[k]] or 0) for k in reasons}
        tot = sum(st.values())
        ex = float(r[col["Instructions Executed"]] or 0)
        e = by_op[op]
        e[0] += tot
        e[1] += ex
        e[2].update(st)
        by_reason.update(st)
        per_ins.append((tot, r[col["Address"]][-5:], src, st))
    lines = [header, "", "total samples %d" % sum(by_reason.values()),
             "by reason: %s" % [(k, v) for k, v in by_reason.most_common() if v], ""]
    for op, (tot, ex, c) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
        lines.append("%-34s %7d  exec %8.1fM  %s" % (op, tot, ex / 1e6, ", ".join("%s %d" % (k[6:], v) for k, v in c.most_common(4) if v)))
    lines.append("")
    for tot, addr, src, st in sorted(per_ins, key=lambda t: -t[0])[:8]:
        lines.append("%d %s %s %s" % (tot, addr, src, [(k[6:], v) for k, v in collections.Counter(st).most_common(3) if v]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
