for r in 1 2 3; do
for lib in /root/repo/libbmx_prev.so ""; do
BMX_LIB=$lib python bench.py --steps 20 --no-e2e --no-cpu --no-verify --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib' or 'new', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done; done
