"""Key metrics of an ncu report (one line per profiled launch).  python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_%'),
        ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'hmma_%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%'),
        ('launch__registers_per_thread', 'regs'), ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_%')]


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].replace('<unnamed>::', '')[:60]
        parts = []
        for key, short in WANT:
            if key in idx:
                parts.append('%s=%s%s' % (short, r[idx[key]], units[idx[key]] if units[idx[key]] not in ('%', '') else ''))
        print(name)
        print('   ' + '  '.join(parts))


if __name__ == '__main__':
    main()
