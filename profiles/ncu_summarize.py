"""Summarise an `ncu --set full` report (read here, on the CPU box) into the table committed under profiles/:

    python profiles/ncu_summarize.py gpurun_out/ncu_r02_targets.ncu-rep > profiles/ncu_targets_r02.txt

One block per profiled launch: duration, DRAM bytes read / written and % of peak, L2 and shared-memory throughput, tensor
and XU (MUFU) pipe activity, issue-slot utilisation, achieved occupancy, registers."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_utchmma_cycles_active.avg.pct_of_peak_sustained_active", "UTCHMMA sub-pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) instructions"),
    ("sm__inst_executed.sum", "instructions executed"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy % (per scheduler)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    kname = col.get("Kernel Name")
    for r in data:
        name = r[kname]
        short = name.split("(")[0].replace("(anonymous namespace)::", "")
        print("== %s   [id %s]" % (short[:110], r[col["ID"]]))
        seen = set()
        for m, label in WANT:
            if m in col and label not in seen and r[col[m]] != "":
                seen.add(label)
                print("   %-44s %14s %s" % (label, r[col[m]], units[col[m]]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
