# BMX_DENSE_LANES: from how many candidate lanes per segment a warp builds all masks right away (dense_tile)
for d in 2 3 4 6; do
  echo "== BMX_DENSE_LANES=$d"
  BMX_DENSE_LANES=$d python profiles/english_bench.py 2>&1 | cut -c1-215 | grep -v "b'e'\|b' '"
  BMX_DENSE_LANES=$d python profiles/short_qgram_bench.py 2>&1 | grep "dna" | cut -c1-120
done
