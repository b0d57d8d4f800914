"""Summarise an .ncu-rep (read here, no GPU): per kernel duration, pipe utilisation, DRAM bytes,
instruction mix per pixel.   python tools/ncu_summary.py REPORT.ncu-rep [pixels]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else 64 * 512 * 512
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.avg", "cyc"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM%"),
        ("smsp__inst_executed.sum", "winst"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel")]
kn = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[kn][:110])
    out = []
    for name, short in want:
        if name in hdr:
            i = hdr.index(name)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            out.append(f"{short}={v:.4g}{(' ' + u) if short in ('rd', 'wr') else ''}")
    print("   " + "  ".join(out))
    try:
        wi = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
        cyc = float(r[hdr.index("sm__cycles_elapsed.avg")].replace(",", ""))
        print(f"   thread-instr/px = {wi * 32 / px:.1f}   clk/px/SM = {cyc * 148 / px:.2f}")
    except Exception:
        pass
if "--mix" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    cur = None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            if cur:
                break
            cur = {"name": r[1], "hdr": None, "ops": collections.Counter(), "samp": collections.Counter()}
            continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r
            ie, si = r.index("Instructions Executed"), r.index("# Samples")
            continue
        op = r[1].split()
        if not op:
            continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        cur["ops"][o] += int(r[ie])
        cur["samp"][o] += int(r[si])
    ts = sum(cur["samp"].values()) or 1
    print("instruction mix of", cur["name"][:80])
    for o, n in cur["ops"].most_common(28):
        print(f"   {o:10s} {n * 32 / px:7.2f} /px   stall samples {100 * cur['samp'][o] / ts:5.1f}%")
