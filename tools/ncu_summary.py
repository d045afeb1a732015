"""Key metrics of an ncu report (`ncu --set full`) as a small text table.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt   (runs `ncu -i ... --page raw --csv`)"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__cycles_elapsed.avg', 'gpc__cycles_elapsed.avg.per_second', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'# {path}')
    for r in rows[2:]:
        print(f"\n## {r[idx['Kernel Name']][:100]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in KEYS:
            if k in idx:
                print(f'{k:75s} {r[idx[k]]:>18s} {units[idx[k]]}')


if __name__ == '__main__':
    main(sys.argv[1])
