set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; cut -c1-300 gpurun_out/bench_r01_n1.json
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_dna_m32.csv $B > gpurun_out/ncu1.log 2>&1
$B --workload aaa_1GiB > gpurun_out/plain_aaa.log 2>&1 && ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_aaa.csv $B --workload aaa_1GiB > gpurun_out/ncu_aaa.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -o gpurun_out/r01_scan_dna -f $B > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:expand_kernel -s 3 -c 1 -o gpurun_out/r01_expand_aaa -f $B --workload aaa_1GiB > gpurun_out/ncu_aaa2.log 2>&1
ls -la gpurun_out/*.ncu-rep
