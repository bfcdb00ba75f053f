"""Condense `ncu -i <rep> --page raw --csv` into the per-kernel table kept under profiles/ (one metric per row, one kernel
launch per column) -- the file bench.py's `roofline.traffic` / `roofline.limiter` are read from.

    ncu -i gpurun_out/r02_oz_gemm_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv "K2 oz_gemm_kernel<6,false,2> (X~^T Y, variables on M)" "K1 oz_gemm_kernel<6,true,2> (Y = X~ A^T)" \
        > profiles/r02_oz_gemm_ncu_full_config3.csv
"""
import csv
import sys

KEEP = [
    "Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__cluster_dim_x",
    "launch__cluster_max_active", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
    "sm__cycles_active.max", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    labels = sys.argv[2:] or ["launch %d" % i for i in range(len(data))]
    out = csv.writer(sys.stdout, lineterminator="\n")
    out.writerow(["metric", "unit"] + labels[:len(data)])
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            out.writerow([name, units[i]] + [d[i] for d in data])


if __name__ == "__main__":
    main()
