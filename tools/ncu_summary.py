"""Turns an ncu report into the markdown + JSON kept under profiles/:
    python tools/ncu_summary.py gpurun_out/r02g.ncu-rep 2368 profiles/r02g_step.md [profiles/r02_stage_counters.json]
Per kernel: duration, warp instructions per window, issue-slot utilisation, resident warps, DRAM bytes per window, the
stall breakdown, the opcode mix and the SASS of the hottest block (instructions executed most often)."""
import csv, io, json, subprocess, sys, collections

rep, W, out_md = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out_json = sys.argv[4] if len(sys.argv) > 4 else None


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, rows = raw[0], raw[1], raw[2:]
col = {k: i for i, k in enumerate(hdr)}


def val(r, k):
    return float(r[col[k]].replace(",", "")) if k in col and r[col[k]] not in ("", "n/a") else float("nan")


md = [f"# ncu summary of `{rep.split('/')[-1]}` ({W} cfg2 windows per launch, `--set full --clock-control none`)\n"]
counters = {}
for r in rows:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("pvio::", "")
    dur = val(r, "gpu__time_duration.sum")
    inst = val(r, "smsp__inst_executed.sum")
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    dram_b = (val(r, "dram__bytes_read.sum") * scale.get(units[col["dram__bytes_read.sum"]], 1.0) +
              val(r, "dram__bytes_write.sum") * scale.get(units[col["dram__bytes_write.sum"]], 1.0))
    dur_us = dur * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]].replace("second", "s").replace("usecond", "us"), 1.0)
    counters[short] = dict(duration_us=dur_us, warp_instructions_per_window=inst / W, dram_bytes_per_window=dram_b / W,
                           issue_active_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           warps_active_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                           registers=val(r, "launch__registers_per_thread"))
    md.append(f"## `{short}`  grid {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} threads, "
              f"{int(val(r, 'launch__registers_per_thread'))} registers\n")
    md.append(f"* duration {dur_us:.1f} us (cold cache, serialised by ncu), {inst / W / 1e3:.1f} K warp instructions / window, "
              f"issue slots {counters[short]['issue_active_pct']:.1f} % busy, resident warps {counters[short]['warps_active_pct']:.1f} % of 64 / SM")
    md.append(f"* DRAM {dram_b / W / 1e3:.1f} KB / window (read + write) = {dram_b / dur_us / 1e3:.0f} GB/s during the launch")
    stalls = sorted(((val(r, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                      .replace("_per_issue_active.ratio", "").replace(".ratio", ""))
                     for k in hdr if "stalled" in k and "ratio" in k and "not_issued" not in k), reverse=True)
    md.append("* stall cycles per issued instruction: " + ", ".join(f"{n} {x:.2f}" for x, n in stalls[:7] if x == x) + "\n")
    # source page of this kernel
    src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--kernel-name", "regex:" + short.split("<")[0]))))
    if len(src) < 3:
        continue
    h2 = src[1]
    iS, iE, iW = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Warp Stall Sampling (All Samples)")
    body = [x for x in src[2:] if len(x) > iE and x[iE].isdigit() and x[0].startswith("0x")]
    seen, uniq = set(), []
    for x in body:                      # the page may list the kernel more than once
        if x[0] in seen:
            break
        seen.add(x[0]); uniq.append(x)
    body = uniq
    ex = collections.Counter()
    for x in body:
        s = x[iS].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op = s.split()[0].rstrip(";").split(".")[0] if s else "?"
        ex[op] += int(x[iE])
    te = sum(ex.values()) or 1
    md.append("opcode mix (warp instructions executed): " + ", ".join(f"{op} {100.0 * n / te:.1f} %" for op, n in ex.most_common(10)) + "\n")
    blocks = []
    for k, x in enumerate(body):
        c = int(x[iE])
        if blocks and blocks[-1][2] == c:
            blocks[-1][1] = k; blocks[-1][3] += c
        else:
            blocks.append([k, k, c, c])
    blocks.sort(key=lambda b: -b[3])
    b = blocks[0]
    md.append(f"hottest block: {b[1] - b[0] + 1} instructions x {b[2] / W:.0f} executions / window = {100.0 * b[3] / te:.0f} % of the kernel's instructions; "
              f"SASS excerpt (stall samples | instruction):\n\n```")
    for x in body[b[0]:min(b[1] + 1, b[0] + 72)]:
        md.append(f"{x[iW]:>5s} | {x[iS].strip()}")
    md.append("```\n")
open(out_md, "w").write("\n".join(md))
if out_json:
    lin = next(v for k, v in counters.items() if k.startswith("lin_obs"))
    sch = next(v for k, v in counters.items() if k.startswith("schur"))
    json.dump({"source": out_md, "windows_per_launch": W,
               "warp_instructions_per_window": lin["warp_instructions_per_window"] + sch["warp_instructions_per_window"],
               "dram_bytes_per_window": lin["dram_bytes_per_window"] + sch["dram_bytes_per_window"],
               "kernels": counters}, open(out_json, "w"), indent=1)
print("wrote", out_md, out_json or "")
