"""Prints the metrics we track from an .ncu-rep (run here, no GPU needed):  python tools/ncu_summary.py rep [--md]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active']
STALLS = 'smsp__average_warps_issue_stalled_'
def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"\n## `{r[idx['Kernel Name']][:110]}`\n\n| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in idx:
                print(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
        st = sorted(((float(r[i] or 0), h) for h, i in idx.items() if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio')), reverse=True)
        for v, h in st[:7]:
            print(f"| stall {h[len(STALLS):-len('_per_issue_active.ratio')]} (per issue) | {v:.3f} | inst |")
main()
