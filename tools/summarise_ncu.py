#!/usr/bin/env python3
"""Turn the two ncu exports of tools/final_evidence.sh into the markdown summaries kept under profiles/.
  python tools/summarise_ncu.py launches gpurun_out/final_launches.csv gpurun_out/final_plain.json > profiles/<name>.md
  python tools/summarise_ncu.py full gpurun_out/final_full_raw.csv > profiles/<name>.md"""
import collections
import csv
import io
import json
import re
import sys


def rows(path):
    txt = [l for l in open(path, errors="replace") if not l.startswith("==")]
    return list(csv.reader(io.StringIO("".join(txt))))


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    return name.strip()


def launches(path, plain_json=None):
    r = rows(path)
    hdr = r[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    per = collections.OrderedDict()
    seq = []
    for row in r[1:]:
        if len(row) <= iv or row[im] != "gpu__time_duration.sum":
            continue
        v = float(row[iv].replace(",", ""))
        if row[iu] in ("ns", "nsecond"):
            v /= 1e3
        elif row[iu] in ("ms", "msecond"):
            v *= 1e3
        seq.append((short(row[ik]), v))
    # one xhe_batch_run of the fast path: ncu serialises kernels in issue order, a run starts with k_layout and its last
    # issued kernel is k_combine_out; take the last complete run of the capture
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_layout")]
    step = []
    for i in reversed(starts):
        ends = [j for j in range(i, len(seq)) if seq[j][0].startswith("k_combine_out")]
        if ends:
            while i > 0 and seq[i - 1][0].startswith(("k_fiat_shamir", "k_sig_hash_prefix")):
                i -= 1                     # the transcript / hash-prefix kernels are issued on their own streams just before k_layout
            step = seq[i:ends[0] + 1]
            break
    for n, v in step:
        per.setdefault(n, [0.0, 0])
        per[n][0] += v; per[n][1] += 1
    total = sum(v for v, _ in per.values())
    print(f"One `xhe_batch_run` (10,000 a1k1 transfers, fast path) under ncu: kernels serialised and cold-cache, so compare SHARES. "
          f"Total {total / 1e3:.2f} ms over {len(step)} launches ({len(seq)} launches in the whole capture).\n")
    print("```")
    for n, (v, c) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f"{n:44s} {v:9.1f} us  {100 * v / total:5.1f}%   x{c}")
    print("```")
    if plain_json:
        d = json.load(open(plain_json))
        print("\nSame command without ncu (CUDA events, isolated serial pass of bench.py, ms): " + json.dumps(d["kernels_ms_per_step_isolated"]))
        print("\nFive-stream timeline of one step of the same run (ms from the start of the step): " + json.dumps(d["timeline_ms_one_step"]))
        print(f"\n`value` of that run: {d['value']:.0f} TX/s, {d['ms_per_step']:.3f} ms per step.")


KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__sass_thread_inst_executed_op_integer_pred_on.sum"]


def full(path):
    r = rows(path)
    hdr, units = r[0], r[1]
    ik = hdr.index("Kernel Name")
    seen = collections.Counter()
    for row in r[2:]:
        if len(row) < len(hdr):
            continue
        n = short(row[ik]); seen[n] += 1
        print(f"## {n}  (launch {seen[n]} of this kernel in the capture)")
        for m in KEEP:
            cols = [i for i, h in enumerate(hdr) if h == m or h.endswith("." + m)]
            if cols:
                i = cols[0]
                print(f"- {m} = {row[i]} {units[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
