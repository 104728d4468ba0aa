"""Extract the per-kernel metrics quoted in DESIGN.md / profiles/*.md from an `ncu --page raw --csv` dump."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__maximum_warps_per_active_cycle_pct',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_sample_count']


def main(path):
    rows = list(csv.reader(open(path)))
    head = rows[0]
    ki = head.index('Kernel Name')
    idx = {w: head.index(w) for w in WANT if w in head}
    best = {}
    for r in rows[2:]:
        name = r[ki].split('(')[0]
        dur = float(r[idx['gpu__time_duration.sum']].replace(',', ''))
        if name not in best or best[name][0] < dur:
            best[name] = (dur, r)
    for name, (dur, r) in sorted(best.items(), key=lambda x: -x[1][0]):
        print('----', name)
        for w in WANT:
            if w in idx:
                print(f"   {w:78s} {r[idx[w]]}")


if __name__ == '__main__':
    main(sys.argv[1])
