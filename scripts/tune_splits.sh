for s2 in 1 2 4; do
  echo "== splits2=$s2"; UCD_SPLITS2=$s2 python scripts/bench_con.py 24 2>&1 | grep "con_sweep\|sweep [12]:" | head -6
done
