"""Summarise an `ncu --set full` capture (raw + source CSV pages) into profiles/: a JSON with the per-launch
numbers bench.py reads (traffic_bytes_per_launch) and a text file with the hottest SASS instructions.
Usage: python tools/ncu_summary.py raw.csv source.csv out.json out_sass.txt "kernel description" "command" [note ...]"""
import collections
import csv
import json
import re
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
         "msecond": 1e3}


def main():
    raw, src, out_json, out_sass, kernel_desc, what = sys.argv[1:7]
    notes = sys.argv[7:]
    rows = list(csv.reader(open(raw)))
    hdr, units, rows = rows[0], rows[1], rows[2:]

    def get(name):
        cols = [i for i, h in enumerate(hdr) if h == name or h.endswith("." + name)]
        for i in cols:
            vals = [r[i].replace(",", "") for r in rows]
            if all(v != "" for v in vals):
                return [float(v) * SCALE.get(units[i], 1) for v in vals]
        return [None] * len(rows)
    keys = {"duration_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum",
            "dram_write_bytes": "dram__bytes_write.sum", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
            "active_warps_per_sm": "sm__warps_active.avg.per_cycle_active",
            "dram_pct_of_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l2_pct_of_peak": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex_pct_of_peak": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm_pct_of_peak": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "tensor_pipe_pct_of_peak": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
            "registers_per_thread": "launch__registers_per_thread", "grid": "launch__grid_size"}
    cols = {k: get(v) for k, v in keys.items()}
    launches = [{k: (round(v[i], 3) if v[i] is not None else None) for k, v in cols.items()} for i in range(len(rows))]
    traffic = [a + b for a, b in zip(cols["dram_read_bytes"], cols["dram_write_bytes"])]

    srows = list(csv.reader(open(src)))
    h = srows[1]
    ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    ins = [(r[ia].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in srows[2:] if len(r) > isamp and r[isamp].isdigit()]
    total = sum(s for _, s, _ in ins) or 1
    classes, ops = collections.Counter(), collections.Counter()
    for t, s, _ in ins:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t)
        if m:
            classes[m.group(2).split(".")[0]] += s
            ops[m.group(2)] += 1
    top = sorted(sorted(enumerate(ins), key=lambda x: -x[1][1])[:30])
    with open(out_sass, "w") as f:
        f.write("# %s\n# hottest SASS instructions by warp-stall samples (ncu --set full, source page): %d samples over %d "
                "instructions\n" % (kernel_desc, total, len(ins)))
        f.write("# samples by opcode of the stalled instruction: %s\n" % ", ".join(
            "%s %.1f%%" % (k, 100 * v / total) for k, v in classes.most_common(12)))
        f.write("# index   share  executed  instruction\n")
        for idx, (t, s, e) in top:
            f.write("%5d  %6.2f%%  %8d  %s\n" % (idx, 100 * s / total, e, t))
        pat = r"(REDG|LDG\.E|LDGSTS|ACQBULK|PREEXIT|STG|UTCHMMA|UBLKCP|LDTM|UTCBAR|UTCATOMSWS|SYNCS|ATOMG|LDGDEPBAR|DEPBAR|BAR|CCTL)"
        f.write("# memory / async / tensor-core opcodes present: %s\n" % ", ".join(
            "%s x%d" % (k, ops[k]) for k in sorted(ops) if re.match(pat, k)))
    json.dump({"what": what, "kernel": kernel_desc, "launches": launches,
               "traffic_bytes_per_launch": int(sum(traffic) / len(traffic)),
               "stall_samples_by_opcode_pct": {k: round(100 * v / total, 1) for k, v in classes.most_common(10)},
               "notes": notes}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
