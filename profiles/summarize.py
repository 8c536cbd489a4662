"""Extracts the judged metrics from ncu reports (gpurun_out/*.ncu-rep) into small text files under profiles/.
usage: python profiles/summarize.py <tag> <report.ncu-rep> [...]"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"gpu__time_duration.sum$|dram__bytes_(read|write)\.sum($|\.per_second)|sm__inst_executed_pipe_fp64.avg.pct|"
                  r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct|launch__(registers_per_thread$|grid_size|"
                  r"block_size|occupancy_limit|shared_mem_per_block_dynamic)|smsp__inst_executed.sum$|sm__inst_executed.avg.per_cycle_elapsed|"
                  r"smsp__issue_active.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|smsp__inst_executed_op_local|"
                  r"dram__throughput.avg.pct|gpu__dram_throughput|sm__throughput.avg.pct|issue_stalled.*ratio|sm__cycles_elapsed.avg|"
                  r"smsp__sass_inst_executed_op_shared|sm__sass_thread_inst_executed_op_d(add|mul|fma)_pred_on.sum$|lts__t_bytes.sum$")


def main():
    tag = sys.argv[1]
    for rep in sys.argv[2:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, unit = rows[0], rows[1]
        name = re.sub(r"\.ncu-rep$", "", rep.split("/")[-1])
        with open(f"profiles/{tag}_{name}.txt", "w") as fo:
            for r in rows[2:]:
                kname = r[head.index("Kernel Name")]
                fo.write(f"# kernel: {kname}\n# source: ncu --set full --clock-control none ({rep})\n")
                for h, u, v in zip(head, unit, r):
                    if KEEP.search(h):
                        fo.write(f"{h:88s} {v:>22s} {u}\n")
        print("wrote", f"profiles/{tag}_{name}.txt")


if __name__ == "__main__":
    main()
