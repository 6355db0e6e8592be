"""Key metrics of every kernel in an .ncu-rep (ncu --set full): duration, DRAM bytes/throughput, pipe utilisation,
occupancy, registers, top stall reasons.  usage: python tools/ncu_summary.py report.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct", "issue_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(rep, "no data")
        continue
    h, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(h, row))
        u = dict(zip(h, units))
        print(f"== {rep.split('/')[-1]} :: {d.get('Kernel Name', '?')[:100]}")
        for key, label in WANT:
            if key in d:
                print(f"   {label:15s} {d[key]:>14s} {u.get(key, '')}")
        stalls = sorted(((float(d[k].replace(',', '')), k[len(STALL_PREFIX):].replace('_per_issue_active.ratio', ''))
                         for k in d if k.startswith(STALL_PREFIX) and k.endswith("ratio") and "not_issued" not in k and d[k]),
                        reverse=True)[:5]
        print("   top stalls     ", ", ".join(f"{n}={v:.2f}" for v, n in stalls))
