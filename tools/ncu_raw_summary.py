"""Print the key raw metrics of every launch in an .ncu-rep (dev tool)."""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'local_load_requests', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in ix:
            print('%-70s %s %s' % (w, r[ix[w]], units[ix[w]]))
