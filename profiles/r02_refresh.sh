# Round-2 evidence in one GPU call:  bash profiles/r02_refresh.sh   (≈ 7 min on one B200; outputs in gpurun_out/).
# gpurun brings back at most 64 MiB: every ncu report is summarised on the box (profiles/summarize_ncu.py) and only two
# reports travel (headline scan, English 'position' scan) for source-level reading at home.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; cut -c1-300 gpurun_out/bench_r02_n1.json
python profiles/call_latency.py > gpurun_out/r02_call_latency.txt 2>&1
python profiles/english_bench.py > gpurun_out/english_bench_r02.txt 2>&1; cat gpurun_out/english_bench_r02.txt
python profiles/verify_ab.py 1,0 0 > gpurun_out/verify_ab_r02.txt 2>&1
python profiles/size_sweep.py > gpurun_out/size_sweep_r02.txt 2>&1
python profiles/quick_bench.py > gpurun_out/quick_bench_r02.txt 2>&1
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-verify --no-configs"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for w in dna_m32_4GiB bytes256_m4_4GiB ascii95_m16_64MiB aaa_1GiB; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 && ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w > gpurun_out/ncu_l_$w.log 2>&1
done
N="ncu --set full --clock-control none --import-source on -c 1 -f"
$N -k regex:scan_kernel -s 3 -o gpurun_out/r02_scan_dna $B > gpurun_out/ncu_f1.log 2>&1
python profiles/one_scan.py english is > gpurun_out/plain_eng.log 2>&1 && $N -k regex:scan_kernel -s 2 -o gpurun_out/r02_scan_eng_is python profiles/one_scan.py english is > gpurun_out/ncu_f2.log 2>&1
$N -k regex:expand_kernel -s 2 -o gpurun_out/r02_expand_eng_is python profiles/one_scan.py english is > gpurun_out/ncu_f3.log 2>&1
$N -k regex:scan_kernel -s 2 -o gpurun_out/r02_scan_eng_occ python profiles/one_scan.py english "occurrences starting from" > gpurun_out/ncu_f4.log 2>&1
$N -k regex:scan_kernel -s 2 -o gpurun_out/r02_scan_eng_position python profiles/one_scan.py english position > gpurun_out/ncu_f5.log 2>&1
$N -k regex:scan_kernel -s 2 -o gpurun_out/r02_scan_dna8 python profiles/one_scan.py dna8 - > gpurun_out/ncu_f6.log 2>&1
$N -k regex:expand_kernel -s 2 -o gpurun_out/r02_expand_dna8 python profiles/one_scan.py dna8 - > gpurun_out/ncu_f7.log 2>&1
for r in gpurun_out/r02_*.ncu-rep; do python profiles/summarize_ncu.py full $r > gpurun_out/$(basename $r .ncu-rep | sed s/r02_/r02_ncu_/).txt 2>&1; done
for w in dna_m32_4GiB bytes256_m4_4GiB ascii95_m16_64MiB aaa_1GiB; do python profiles/summarize_ncu.py launches gpurun_out/r02_launches_$w.csv > gpurun_out/r02_launches_$w.txt 2>&1; done
for r in gpurun_out/r02_*.ncu-rep; do case $r in *r02_scan_dna.ncu-rep|*r02_scan_eng_position.ncu-rep) ;; *) rm -f $r;; esac; done
du -sh gpurun_out
