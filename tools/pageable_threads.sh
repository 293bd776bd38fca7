for t in 8 16 24 32; do
  echo -n "copy threads $t: "
  FANLIN_COPY_THREADS=$t python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print(round(e['value'],1), round(e['pageable']['value'],1), round(e['pageable']['frac_of_pinned'],3))"
done
nproc
