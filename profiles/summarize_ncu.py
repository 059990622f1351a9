"""ncu -i X.ncu-rep --page raw --csv  ->  compact per-launch summary (the metrics DESIGN.md / bench.py quote)."""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum',
    'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print('== %s  grid %s block %s' % (d[idx['Kernel Name']][:60], d[idx['Grid Size']], d[idx['Block Size']]))
        for w in WANT:
            if w in idx:
                print('   %-72s %16s %s' % (w, d[idx[w]], units[idx[w]]))


if __name__ == '__main__':
    main(sys.argv[1])
