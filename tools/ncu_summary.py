"""Key counters of every launch in an .ncu-rep (read here on the CPU box: ncu -i ... --page raw --csv), as committed under profiles/.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("dram__cycles_active.min.pct_of_peak_sustained_elapsed", "DRAM busiest/idlest channel: min active"),
    ("dram__cycles_active.max.pct_of_peak_sustained_elapsed", "  max active"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (achieved occupancy)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks)"),
]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("== %s" % rep)
    for r in rows[2:]:
        print("-- %s" % r[hdr.index("Kernel Name")])
        for key, label in WANT:
            if key in hdr and r[hdr.index(key)] != "":
                print("   %-48s %s %s" % (label, r[hdr.index(key)], units[hdr.index(key)]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        top = sorted(stalls, reverse=True)[:5]
        if top:
            print("   %-48s %s" % ("top stall reasons (share of samples)", ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in top)))
