"""Extract the judged metrics from an .ncu-rep into a small text file (run in the authoring container).
usage: python tools/ncu_extract.py <rep> <out.txt> "<title>" """
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
with open(out, "w") as f:
    f.write(f"# {title}\n# source: {rep} (ncu --set full --clock-control none; cold caches, serialised launches)\n")
    for r in rows[2:]:
        f.write("---\n")
        for i, name in enumerate(h):
            if name in keep and r[i] not in ("", "n/a"):
                f.write(f"{name} = {r[i]} {rows[1][i]}\n")
print(open(out).read()[:1500])
