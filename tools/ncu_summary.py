"""Summarise an ncu report (one line block per kernel launch) from its raw page.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex_%'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_wavefronts_%'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'wf_shared'),
    ('l1tex__data_pipe_lsu_wavefronts.sum', 'wf_all'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pipe_%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_%'),
    ('launch__registers_per_thread', 'regs'), ('launch__occupancy_limit_registers', 'lim_regs'),
    ('launch__occupancy_limit_shared_mem', 'lim_smem'), ('launch__occupancy_limit_warps', 'lim_warps'),
    ('launch__shared_mem_per_block_dynamic', 'smem/blk'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_%'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'gld_sectors'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'gld_requests'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'gst_sectors'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'gst_requests'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_%'),
]
STALL = 'smsp__average_warps_issue_stalled_'


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            print("== %s :: %s grid=%s block=%s" % (path, r[idx['Kernel Name']][:110], r[idx['Grid Size']], r[idx['Block Size']]))
            for key, label in WANT:
                if key in idx:
                    print("   %-22s %s %s" % (label, r[idx[key]], units[idx[key]]))
            stalls = []
            for h, i in idx.items():
                if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                    try:
                        stalls.append((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("   stalls (warps per issue-active cycle): " + ", ".join("%s=%.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == '__main__':
    main()
