#!/usr/bin/env python
"""Print the metrics the roofline discussion uses from an .ncu-rep (run in the build container: no GPU needed)."""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__pipe_tensor_cycles_active.*pct|"
                 r"sm__pipe_tensor_subpipe_imma_cycles_active.*pct|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
                 r"launch__shared_mem_per_block_dynamic|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate.pct|"
                 r"lts__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__data_bank_(reads|writes).avg.pct_of_peak_sustained_elapsed|"
                 r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|smsp__inst_executed.sum|sm__cycles_elapsed.avg|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum|sm__inst_executed_pipe_uniform.sum|lts__t_bytes.sum)$")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("==", name[:100])
        for h, u, v in zip(hdr, units, r):
            if PAT.match(h):
                print("   %-75s %-8s %s" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1])
