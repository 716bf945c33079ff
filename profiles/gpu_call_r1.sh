# Round-1 profiling call (run under gpurun): launch list + ncu --set full of the hot kernels, summaries exported
# as CSV on the box so they survive the 64 MiB gpurun_out limit.
mkdir -p gpurun_out
RAW='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__pipe_tensor_subpipe|sm__throughput.avg.pct|gpu__dram_throughput|lts__throughput|lts__t_bytes.sum|l1tex__throughput|sm__warps_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__cycles_active.avg|smsp__cycles_active.avg|sm__inst_executed_pipe_tensor|smsp__warp_issue_stalled|launch__occupancy_limit|sm__cycles_elapsed.max|lts__t_sector_hit_rate|lts__t_sectors_srcunit_tex_op_read.sum'
python profiles/time_kernels.py > gpurun_out/time_kernels.log 2>&1
python profiles/timeline_tblock.py > gpurun_out/timeline_tblock.log 2>&1
python profiles/run_one.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python profiles/run_one.py > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python profiles/run_one.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'tblock_kernel|attn_kernel' -c 4 -o gpurun_out/prof_tblock_attn python profiles/run_one.py > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
python profiles/run_one.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --profile-from-start off -k regex:conv_gemm_kernel -c 98 -o gpurun_out/prof_conv python profiles/run_one.py > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
for f in prof_tblock_attn prof_conv; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_tblock_attn.ncu-rep --page details --csv > gpurun_out/prof_tblock_attn.details.csv 2>/dev/null
ncu -i gpurun_out/prof_tblock_attn.ncu-rep --page source --csv > gpurun_out/prof_tblock_attn.source.csv 2>/dev/null
ls -la gpurun_out
# keep the return under 64 MiB: drop the biggest raw reports first (their CSV exports stay)
while [ "$(du -sm gpurun_out | cut -f1)" -ge 60 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"; echo "dropped $big"
done
du -sm gpurun_out
